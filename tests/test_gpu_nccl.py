"""-m gpu, needs >= 2 GPUs: the concept-parallel path on real hardware — tmx_blend_partial_fwd -> NCCL all-reduce ->
tmx_blend_finish_fwd and the sharded sampler — launched under torch.distributed.run (see tests/nccl_worker.py for the
checks and tolerances).  Skipped on a 1-GPU box; the gloo tests of tests/test_dist_gloo.py cover the host logic there."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_nccl_matches_single_rank(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "nccl_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    rep = json.loads(out.read_text())
    print("NCCL parity:", rep)
    assert rep["kernel_ranks_bit_identical"] and rep["kernel_max_abs_vs_single_gpu"] <= 2e-4
    assert rep["sampler_custom_ranks_bit_identical"] and rep["sampler_lora_ranks_bit_identical"]
