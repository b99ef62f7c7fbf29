"""Host logic of uorf_b200.optim.Adam and modules.invert_pose that needs no GPU: on CPU tensors both hand over to torch's own
implementation (the kernels are CUDA-only), with torch.optim.Adam's state layout."""
import torch

from uorf_b200.optim import Adam, _CHUNK_DTYPE
from uorf_b200.modules import invert_pose


def test_adam_on_cpu_is_torch_adam():
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = Adam(a, lr=1e-2), torch.optim.Adam(b, lr=1e-2)
    for _ in range(5):
        for p, q in zip(a, b):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert oa.kernel_steps == 0                      # nothing ran on the CUDA kernel
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"] == sb["param_groups"]
    for k in sb["state"]:
        for name in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(torch.as_tensor(sa["state"][k][name]), torch.as_tensor(sb["state"][k][name]))
    ob.load_state_dict(sa)                           # same layout both ways
    oa.load_state_dict(sb)


def test_adam_closure_and_chunk_struct():
    p = torch.nn.Parameter(torch.ones(3))
    o = Adam([p], lr=0.1)

    def closure():
        o.zero_grad()
        loss = (p * p).sum()
        loss.backward()
        return loss

    loss = o.step(closure)
    assert float(loss) == 3.0 and float(p[0]) < 1.0
    # struct uorf_adam_chunk (include/uorf_b200.h): five pointers, two int32
    assert _CHUNK_DTYPE.itemsize == 48 and _CHUNK_DTYPE.names == ("param", "grad", "exp_avg", "exp_avg_sq", "step", "n", "reserved")


def test_invert_pose_on_cpu_is_torch_inverse():
    m = torch.eye(4)[None].repeat(3, 1, 1) + 0.1 * torch.randn(3, 4, 4)
    assert torch.equal(invert_pose(m), m.inverse())
    d = torch.randn(2, 3, 3, dtype=torch.float64)
    assert torch.equal(invert_pose(d), d.inverse())
