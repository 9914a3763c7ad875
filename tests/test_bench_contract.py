"""bench.py contract pieces that run without a GPU: the reference arm's JSON line (with the CPU forward stubbed so the
test is fast), exactly-one-line stdout protection, and the ncu-traffic lookup used for `roofline.traffic`."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_reference_arm_line(monkeypatch, capsys):
    b = _load_bench()
    import oracle.cpu_baseline as cb
    monkeypatch.setattr(cb.CpuReference, "__init__", lambda self, res=1024: setattr(self, "threads", 3) or None)
    monkeypatch.setattr(cb.CpuReference, "forward_seconds", lambda self: 2.0)
    monkeypatch.setattr(cb.CpuReference, "sample_description", lambda self: "stub")
    monkeypatch.setenv("RANK", "0")
    args = b.parse_args(["--impl", "reference", "--steps", "2", "--warmup", "1", "--gpus", "4"]) if hasattr(b, "parse_args") else None
    if args is None:
        monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1", "--gpus", "4"])
        args = b.parse()
    assert b.run_reference(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 4 and line["steps"] == 2 and line["gpu_launches"] == 0
    assert abs(line["value"] - 1.0 / (2.0 * 242)) < 1e-12                      # 242 sample-forwards per 50-step image
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 3
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["baseline_config"] == "configs[1]" and line["config"]["group_size"] == 2 and line["config"]["image_groups"] == 2
    assert line["scaling"] == "weak" and line["config"]["sample_forwards_per_image"] == 242
    # other ranks of a torchrun launch print nothing and exit 0
    monkeypatch.setenv("RANK", "1")
    assert b.run_reference(args) == 0 and capsys.readouterr().out == ""


def test_ncu_traffic_lookup():
    b = _load_bench()
    got = b._ncu_attention_traffic({"Nq4096_Nk4096_H10": 10, "Nq1024_Nk1024_H20": 60})
    assert got["traffic"] is not None and 1e7 < got["traffic"] < 1e8 and "ncu --set full" in got["traffic_source"]
    assert b._ncu_attention_traffic({"Nq4096_Nk77_H10": 1}) == {"traffic": None}          # no capture for that shape


def test_stdout_carries_only_the_json_line():
    """Library chatter on fd 1 (NCCL's version banner under torchrun) must not precede the result line."""
    code = ("import os, sys; sys.argv=['bench.py']; sys.path.insert(0, %r); import importlib.util as u; "
            "s=u.spec_from_file_location('b', %r); b=u.module_from_spec(s); s.loader.exec_module(b); "
            "b._protect_stdout(); os.write(1, b'NCCL version banner\\n'); print('python chatter'); b._emit({'ok': 1})"
            % (ROOT, os.path.join(ROOT, "bench.py")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version banner" in r.stderr and "python chatter" in r.stderr


def test_workload_specs(monkeypatch):
    """Group-size policy and the forward counts of the three runnable BASELINE configs."""
    b = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    args = b.parse()
    for world, want in ((1, (1, 1)), (2, (2, 1)), (4, (2, 2)), (8, (2, 4))):
        sp = b.resolve_spec(args, world)
        assert (sp["group"], sp["n_groups"]) == want and b.scaling_label(sp) == ("weak" if want[1] > 1 else "strong")
    sp2 = b.resolve_spec(args, 4, 2)
    assert (sp2["variant"], sp2["group"], sp2["n_groups"]) == ("lora", 4, 1) and b.forwards_per_image(sp2) == 224
    sp3 = b.resolve_spec(args, 8, 3)
    assert (sp3["concept_num"], sp3["image_batch"], sp3["group"], sp3["n_groups"]) == (9, 4, 8, 1)
    assert b.forwards_per_image(b.resolve_spec(args, 1)) == 242
    monkeypatch.setattr(sys, "argv", ["bench.py", "--group-size", "4", "--gpus", "8"])
    sp = b.resolve_spec(b.parse(), 8)
    assert (sp["group"], sp["n_groups"]) == (4, 2)
