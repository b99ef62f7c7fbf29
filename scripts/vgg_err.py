import torch, sys, copy, time
sys.path.insert(0,'.')
from uorf_b200.modules import PerceptualVGG
from uorf_b200.nogan_model import perceptual_net
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(11)
dev=torch.device('cuda:0')
net = perceptual_net(None, allow_random_init=True).to(dev)
fast = PerceptualVGG(net)
net64 = copy.deepcopy(net).double()
img = torch.rand(4, 3, 64, 64, device=dev) * 2 - 1
tgt = torch.rand(4, 3, 64, 64, device=dev) * 2 - 1
res={}
for name,f,dt in (("f64",net64,torch.float64),("cudnn",net,torch.float32),("x3",fast,torch.float32)):
    a=img.to(dt).clone().requires_grad_(True)
    fa=f(a)
    with torch.no_grad(): fb=f(tgt.to(dt))
    loss=0.006*torch.nn.functional.mse_loss(fa,fb); loss.backward()
    res[name]=(fa.detach().double(), loss.item(), a.grad.double())
for n in ("cudnn","x3"):
    print(n, "feat err", (res[n][0]-res["f64"][0]).abs().max().item(), "of", res["f64"][0].abs().max().item(),
          "loss rel", abs(res[n][1]-res["f64"][1])/res["f64"][1], "grad err", (res[n][2]-res["f64"][2]).abs().max().item(), "of", res["f64"][2].abs().max().item())
# timing
for name,f in (("cudnn",net),("x3",fast)):
    x8=torch.rand(8,3,64,64,device=dev)
    for _ in range(3): f(x8)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(20): f(x8)
    torch.cuda.synchronize(); print(name,"fwd batch 8 ms",(time.time()-t)/20*1e3)
