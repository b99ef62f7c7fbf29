"""CPU emulation: which rounding of a 16-bit decoder backward moves the 20-step training trajectory off the reference's?
Runs the oracle's training loop on tests/golden/train_trajectory.npz with the decoder evaluated under emulated roundings
(straight-through, so gradients flow to the fp32 weights):  W: weights -> fp16;  A: activations -> fp16 after every ReLU;
G: activation gradients -> fp16 under a power-of-two loss scale.  Prints the max relative deviation of the loss curve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import torch.nn.functional as F
from conftest import load_golden
from oracle import uorf_oracle as O

MODE = {"W": False, "A": False, "G": False, "W2": False}


def r16(x):
    return x + (x.half().float() - x).detach()


def r16x2(x):       # two-term split: hi + lo
    hi = x.half().float()
    lo = (x - hi).half().float()
    return x + ((hi + lo) - x).detach()


class GradRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        m = g.abs().max().clamp(min=1e-30)
        s = 2.0 ** torch.floor(torch.log2(16384.0 / m))
        return (g * s).half().float() / s


def lin(p, name, x):
    w = p[name + ".weight"]
    if MODE["W"]:
        w = r16(w)
    if MODE["W2"]:
        w = r16x2(w)
    y = F.linear(x, w, p[name + ".bias"])
    if MODE["G"]:
        y = GradRound.apply(y)
    return y


def trunk(p, pre, inp, n_layers):
    t = inp
    for i in range(n_layers):
        t = F.relu(lin(p, f"{pre}_before.{2 * i}", t))
        if MODE["A"]:
            t = r16(t)
    t = torch.cat([inp, t], dim=1)
    for i in range(n_layers):
        t = F.relu(lin(p, f"{pre}_after.{2 * i}", t))
        if MODE["A"]:
            t = r16(t)
    return t


O._lin, O._mlp_trunk = lin, trunk


def run(tag, **mode):
    MODE.update({k: False for k in MODE})
    MODE.update(mode)
    g = load_golden("train_trajectory")
    spec = O.FrustumSpec([8, 8, 8], render_size=8)
    P = {pre: {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["init"][pre].items()} for pre in ("enc", "sa", "dec")}
    params = [v for pre in ("enc", "sa", "dec") for v in P[pre].values() if v.requires_grad]
    optim = torch.optim.Adam(params, lr=3e-4)
    losses = []
    for i in range(20):
        for grp in optim.param_groups:
            grp["lr"] = float(g["lr"][i])
        loss, *_ = O.train_forward(P["enc"], P["sa"], P["dec"], g["img"], g["cam2world"], g["azi_rot"], spec, num_slots=3, stage="coarse",
                                   input_size=32, supervision_size=8, noise_fg=g["noise_fg"][i], noise_bg=g["noise_bg"][i])
        optim.zero_grad(); loss.backward(); optim.step(); losses.append(loss.item())
    ref = g["loss_recon"].double()
    rel = ((torch.tensor(losses, dtype=torch.float64) - ref).abs() / ref)
    print(f"{tag:28s} max rel dev {rel.max().item():.2e}  at step {int(rel.argmax())}  final {losses[-1]:.5f} (ref {ref[-1]:.5f})")


if __name__ == "__main__":
    run("fp32")
    run("weights fp16", W=True)
    run("weights fp16x2", W2=True)
    run("activations fp16", A=True)
    run("act grads fp16", G=True)
    run("W+A+G (1-pass backward)", W=True, A=True, G=True)
    run("W2+A+G", W2=True, A=True, G=True)
    run("W2 + G", W2=True, G=True)
