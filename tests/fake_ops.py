"""TEST INFRASTRUCTURE ONLY: plain-PyTorch stand-ins for ``tweediemix_b200.ops`` so that the HOST
logic of the product (U-Net wiring, hook routing, sampler phases, concept-parallel row assignment)
can be exercised on the CPU-only build box against the oracle.  The product never imports this;
``install(monkeypatch)`` swaps the functions in for one test.  The arithmetic of the blend functions
is the oracle's (``oracle/step_math.py``), i.e. the checker, not the thing shipped.
"""
import torch
import torch.nn.functional as F

from oracle import step_math as sm


def group_norm(x, gamma, beta, groups, eps, *, silu=False, add=None, out=None, x2=None):
    if x2 is not None:
        x = torch.cat([x, x2], dim=1)
    xf = x.float()
    if add is not None:
        xf = xf + add[:, :, None, None]
    y = F.group_norm(xf, groups, gamma, beta, eps)
    y = (F.silu(y) if silu else y).to(x.dtype)
    y = y.contiguous(memory_format=torch.channels_last) if not x.is_contiguous() else y
    if out is not None:
        out.copy_(y)
        return out
    return y


def residual_add(a, b, inv_scale=1.0, out=None):
    y = ((a.float() + b.float()) * inv_scale).to(a.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def bias_residual_add(a, bias, b=None, inv_scale=1.0, out=None):
    shape = (1, -1, 1, 1) if a.dim() == 4 else (-1,)
    y = a.float() + bias.reshape(shape)
    if b is not None:
        y = y + b.float()
    y = (y * inv_scale).to(a.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def residual_add_layer_norm(a, b, gamma, beta, eps, h_out=None, n_out=None):
    h = (a.float() + b.float()).to(a.dtype)
    n = F.layer_norm(h.float(), (h.shape[-1],), gamma, beta, eps).to(a.dtype)
    if h_out is not None:
        h_out.copy_(h)
        h = h_out
    return h, n


def layer_norm(x, gamma, beta, eps, out=None):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).to(x.dtype)


def geglu(x, out=None):
    h, g = x.chunk(2, dim=-1)
    return h * F.gelu(g)


def attention(q, k, v, heads, scale=None, out=None):
    B, Nq, HD = q.shape
    D = HD // heads
    scale = D ** -0.5 if scale is None else scale
    qh = q.reshape(B, Nq, heads, D).permute(0, 2, 1, 3).float()
    kh = k.reshape(B, -1, heads, D).permute(0, 2, 1, 3).float()
    vh = v.reshape(B, -1, heads, D).permute(0, 2, 1, 3).float()
    p = (qh @ kh.transpose(-1, -2) * scale).softmax(dim=-1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, Nq, HD).to(q.dtype)


def routed_linear(x, weights=None, lora_down=None, lora_up=None, *, nseg=1, out=None):
    B, M, Kin = x.shape
    if weights is not None:
        y = torch.stack([x[b].float() @ weights[b].float().t() for b in range(B)]).to(x.dtype)
        if out is not None:
            out.copy_(y)
        else:
            out = y
    if lora_down is not None:
        Nout = out.shape[-1]
        seg = Nout // nseg
        for b in range(B):
            if lora_down[b] is None:
                continue
            r = lora_down[b].shape[0] // nseg
            t = x[b].float() @ lora_down[b].float().t()                       # [M, nseg*r]
            for s_ in range(nseg):
                up = lora_up[b][s_ * seg:(s_ + 1) * seg].float()              # [seg, r]
                out[b, :, s_ * seg:(s_ + 1) * seg] += (t[:, s_ * r:(s_ + 1) * r] @ up.t()).to(out.dtype)
    return out


def layout_supported(*ts):
    return all(t.dim() == 4 and t.shape[1] % 8 == 0 for t in ts)


def cat_channels(a, b, out=None):
    y = torch.cat([a, b], dim=1)
    return y if out is None else out.copy_(y)


def upsample_nearest2x(x, out=None):
    y = torch.nn.functional.interpolate(x, scale_factor=2.0, mode="nearest")
    return y if out is None else out.copy_(y)


def cat_free_supported(a, b):
    return a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0


def linear_supported(x, w):
    from tweediemix_b200 import ops
    return ops.GEMM_IMPL != "cublas" and x.shape[-1] % 64 == 0 and w.shape[0] % 8 == 0


def linear(x, w, bias=None, *, residual=None, geglu=False, lora_tail=None, out=None):
    N = w.shape[0]
    y = x.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if lora_tail is not None:                                    # the K = 16 tail step: t[:, :16] @ up_ext[b][:, :16].T per batch row
        t, ups, rpb = lora_tail
        y2 = y.reshape(-1, N)
        for b, u in enumerate(ups):
            if u is not None:
                y2[b * rpb:(b + 1) * rpb] += t[b * rpb:(b + 1) * rpb, :16].float() @ u[:, :16].float().t()
        y = y2.reshape(y.shape)
    if geglu:                                                    # rows interleaved in blocks of 32 value | 32 gate
        z = y.reshape(*y.shape[:-1], N // 64, 2, 32)
        y = (z[..., 0, :] * F.gelu(z[..., 1, :])).reshape(*y.shape[:-1], N // 2)
    if residual is not None:
        y = y + residual.float()
    y = y.to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def lora_t(x, downs, sr):
    B, M, K = x.shape
    t = torch.zeros(B * M, 64, dtype=x.dtype, device=x.device)
    for b, d in enumerate(downs):
        if d is not None:
            t[b * M:(b + 1) * M, :sr] = (x[b].float() @ d.float().t()).to(x.dtype)
    return t


def vpred_cfg_ddim(x, v_uncond, v_cond, a_t, a_next, g, *, out=None, x0_out=None, ref_rounding=False):
    xf, u, c = x.float(), v_uncond.float(), v_cond.float()
    v = u + g * (c - u)
    sa, sb, sna, snb = a_t ** 0.5, (1 - a_t) ** 0.5, a_next ** 0.5, (1 - a_next) ** 0.5
    eps = sa * v + sb * xf
    x0 = sa * xf - sb * v
    res = (sna * x0 + snb * eps).to(x.dtype)
    if x0_out is not None:
        x0_out.copy_(x0.to(x.dtype))
    if out is not None:
        out.copy_(res)
        return out
    return res


def frame_inject(y, groups, frames, interp=1.0, *, ref_rounding=False):
    v = y.reshape(groups, frames, -1)
    first = v[:, :1].float()
    v[:, 1:] = (interp * first + (1.0 - interp) * v[:, 1:].float()).to(y.dtype) if interp != 1.0 else v[:, :1].expand_as(v[:, 1:])
    return y


def _mw(masks, weights, K, like):
    m = masks if masks is not None else torch.ones(K, 1, 1, 1, device=like.device)
    if weights is not None:
        m = m * torch.tensor(weights, dtype=torch.float32, device=like.device).reshape(K, 1, 1, 1)
    return m


def tweedie_blend_ddim(x, eps, masks, a_t, a_next, g, *, is_last=False, weights=None, out=None, x0_out=None,
                       ref_rounding=False):
    imgs, C, H, W = x.shape
    e = eps.reshape(imgs, -1, C, H, W).float()                 # [imgs, K+1, C, H, W], image-major like the real op
    K = e.shape[1] - 1
    res, x0 = zip(*[sm.fused_step(x[i:i + 1], e[i], _mw(masks, weights, K, x), a_t, a_next, g, is_last=is_last) for i in range(imgs)])
    res, x0 = torch.cat(res), torch.cat(x0)
    if x0_out is not None:
        x0_out.copy_(x0)
    if out is not None:
        out.copy_(res)
        return out
    return res


def blend_partial(eps_rows, masks, row_ids, acc, imgs=1, *, K=None, weights=None):
    K = masks.shape[0] if K is None else K
    mw = _mw(masks, weights, K, acc)
    acc.zero_()
    for j, r in enumerate(row_ids):
        e = eps_rows.reshape(len(row_ids), *acc.shape[2:])[j].float()
        if r == 0:
            acc[0, 1] = e
        else:
            acc[0, 0] += mw[r - 1] * e
    return acc


def blend_finish(x, acc, masks, a_t, a_next, g, *, is_last=False, out=None, x0_out=None, K=None, weights=None):
    K = masks.shape[0] if K is None else K
    M = _mw(masks, weights, K, x).sum(dim=0, keepdim=True).expand(1, 1, *x.shape[2:])
    res, x0 = zip(*[sm.blend_finish(x[i:i + 1], acc[i:i + 1, 0], acc[i:i + 1, 1], M, a_t, a_next, g, is_last=is_last) for i in range(x.shape[0])])
    res, x0 = torch.cat(res), torch.cat(x0)
    if x0_out is not None:
        x0_out.copy_(x0)
    if out is not None:
        out.copy_(res)
        return out
    return res


NAMES = ("group_norm", "layer_norm", "residual_add", "bias_residual_add", "residual_add_layer_norm", "geglu", "attention",
         "tweedie_blend_ddim", "blend_partial", "blend_finish", "routed_linear", "linear", "lora_t", "linear_supported",
         "vpred_cfg_ddim", "frame_inject", "cat_free_supported", "layout_supported", "cat_channels", "upsample_nearest2x")


def install(monkeypatch):
    from tweediemix_b200 import ops
    for name in NAMES:
        monkeypatch.setattr(ops, name, globals()[name])
