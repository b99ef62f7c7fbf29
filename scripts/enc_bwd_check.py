"""per-kernel check of the Encoder gradient kernels against aten.convolution_backward (fp32, cuDNN) - prints relative errors"""
import torch, torch.nn.functional as F
from uorf_b200._lib import lib, check

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
S = torch.cuda.current_stream().cuda_stream


def run(ca, cb, up, c_out, h, w, stride):
    torch.manual_seed(ca * 7 + h + stride)
    a = torch.randn(ca, h // 2, w // 2, device=dev) if up else torch.randn(ca, h, w, device=dev)
    b = torch.randn(cb, h, w, device=dev) if cb else None
    xa = F.interpolate(a[None], scale_factor=2, mode="bilinear", align_corners=False)[0] if up else a
    x = torch.cat([xa, b]) if cb else xa
    wt = torch.randn(c_out, ca + cb, 3, 3, device=dev) * 0.05
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    kept = torch.randn(c_out, ho, wo, device=dev)
    g = torch.randn(c_out, ho, wo, device=dev)
    gi, gw, gb = torch.ops.aten.convolution_backward((g * (kept > 0))[None], x[None], wt, [c_out], [stride, stride], [1, 1], [1, 1], False, [0, 0], 1,
                                                     [True, True, True])
    # float64 reference
    gi64, gw64, gb64 = torch.ops.aten.convolution_backward((g * (kept > 0))[None].double(), x[None].double(), wt.double(), [c_out], [stride, stride], [1, 1],
                                                           [1, 1], False, [0, 0], 1, [True, True, True])
    cin = ca + cb
    ws = torch.empty(lib().uorf_encoder_wgrad_workspace_floats(cin, c_out, h, w, stride), device=dev)
    mw, mb = torch.full_like(wt, float("nan")), torch.full((c_out,), float("nan"), device=dev)
    check(lib().uorf_encoder_conv3x3_wgrad(a.data_ptr(), ca, int(up), b.data_ptr() if cb else None, cb, g.data_ptr(), kept.data_ptr(), c_out, h, w, stride,
                                           ws.data_ptr(), mw.data_ptr(), mb.data_ptr(), S), "wgrad")
    rel = lambda m, r: ((m.double() - r).abs().max() / r.abs().max()).item()
    out = f"ca {ca:3d} cb {cb:3d} up {int(up)} co {c_out:3d} {h:3d}x{w:<3d} s{stride}: wgrad {rel(mw, gw64):.2e} (cudnn {rel(gw, gw64):.2e}) bias {rel(mb, gb64):.2e}"
    if cin % 4 == 0:
        wg = wt.permute(0, 2, 3, 1).contiguous()
        mi = torch.full((cin, h, w), float("nan"), device=dev)
        check(lib().uorf_encoder_conv3x3_dgrad(g.data_ptr(), kept.data_ptr(), c_out, wg.data_ptr(), cin, h, w, stride, 0, mi.data_ptr(), S), "dgrad")
        out += f" dgrad {rel(mi, gi64[0]):.2e} (cudnn {rel(gi[0], gi64[0]):.2e})"
        base = torch.randn_like(mi)
        acc = base.clone()
        check(lib().uorf_encoder_conv3x3_dgrad(g.data_ptr(), kept.data_ptr(), c_out, wg.data_ptr(), cin, h, w, stride, 1, acc.data_ptr(), S), "dgrad")
        out += f" acc {rel(acc - base, gi64[0]):.2e}"
    print(out, flush=True)


for cfg in [(3, 4, False, 64, 64, 64, 1), (3, 4, False, 64, 128, 128, 1), (64, 0, False, 64, 128, 128, 2), (64, 0, False, 64, 64, 64, 2),
            (64, 0, False, 64, 32, 32, 2), (64, 0, False, 64, 16, 16, 1), (64, 64, True, 64, 32, 32, 1), (64, 64, True, 64, 64, 64, 1),
            (64, 64, True, 64, 128, 128, 1), (40, 40, True, 40, 24, 24, 1), (40, 0, False, 40, 12, 12, 2), (64, 0, False, 64, 256, 256, 2),
            (3, 4, False, 64, 256, 256, 1), (64, 0, False, 64, 20, 28, 2), (64, 64, True, 64, 20, 28, 1)]:
    run(*cfg)
# upsample backward
for c, h, w in [(64, 16, 16), (40, 6, 6), (64, 32, 32), (3, 1, 1), (5, 7, 3)]:
    low = torch.randn(1, c, h, w, device=dev, dtype=torch.float64, requires_grad=True)
    up = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=False)
    g = torch.randn_like(up)
    up.backward(g)
    out = torch.full((c, h, w), float("nan"), device=dev)
    gf = g[0].float().contiguous()
    check(lib().uorf_encoder_upsample2x_bwd(gf.data_ptr(), c, h, w, out.data_ptr(), S), "up_bwd")
    print(f"upsample bwd {c}x{h}x{w}: {((out.double() - low.grad[0]).abs().max() / low.grad.abs().max()).item():.2e}", flush=True)
