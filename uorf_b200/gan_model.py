"""Adversarial training driver: the B200 mirror of the reference's uorfGanModel (models/uorf_gan_model.py:15-375), the model
the shipped Room-Diverse recipe trains through (scripts/train_room_diverse.sh, train_with_gan.py:36-56).

It is the no-GAN driver plus a discriminator: the generator step adds `weight_gan * softplus(-D(x_recon))` to the loss
(:188-189), and `forward_disc(it)` renders the scene again and takes one discriminator step (logistic loss, R1 penalty every
32nd call, :230-294).  The render inside both steps is the same hot path (Encoder -> SlotAttention -> sampling -> fused
decoder -> compositing) as `uorfNoGanModel`; the discriminator is a small cuDNN network and stays torch, like the
perceptual net.

Differences from the reference, none of which changes a result:
  * `forward_disc` renders the fake image under `torch.no_grad()`.  The reference builds the generator's autograd graph,
    back-propagates d_loss into the generator and throws those gradients away at the next `zero_grad` (:235-236); only the
    discriminator's own gradients are used, and they do not depend on the generator's graph.
  * `conv2d_gradfix` (models/op/conv2d_gradfix.py) is a cuDNN-flag wrapper around conv2d with a double-backward path;
    torch's own conv2d double backward computes the same R1 gradient.

Same class names, constructor arguments and state-dict keys (`convs.{i}.conv{1,2}.{0|1}.weight`, `final_linear.{0,1}.*`)
as models/model.py:352-549, so `{suffix}_net_Disc.pth` files load unchanged.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F
from torch import autograd, nn, optim

from .nogan_model import default_train_options, exp_decay_schedule_with_warmup, uorfNoGanModel


def default_gan_options(**over):
    """uorfGanModel.modify_commandline_options (uorf_gan_model.py:28-60): the no-GAN options with the GAN defaults
    (percept_in 20, no_locality_epoch 60, input_size 128, coarse_epoch 120) plus d_lr, weight_gan, gan_in, ndf, weight_r1,
    gan_train_epoch."""
    opt = default_train_options(percept_in=20, no_locality_epoch=60, input_size=128, coarse_epoch=120, d_lr=1e-3,
                                weight_gan=0.01, gan_in=5, ndf=64, weight_r1=10.0, gan_train_epoch=20)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


# ---- losses, reference models/model.py:327-349 -----------------------------------------------------------
def toggle_grad(model, requires_grad):
    for p in model.parameters():
        p.requires_grad_(requires_grad)


def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean(), F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    grad_real, = autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


# ---- discriminator, reference models/model.py:362-549 ----------------------------------------------------
class EqualConv2d(nn.Module):
    """conv with N(0,1) weights scaled by 1/sqrt(fan_in) at run time (model.py:362-397)"""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride, self.padding = stride, padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return F.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)


class EqualLinear(nn.Module):
    """model.py:400-435; activation 'fused_lrelu' drops the bias and applies leaky_relu(0.2) * 1.4, as the reference does"""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            return F.leaky_relu(F.linear(input, self.weight * self.scale), 0.2) * 1.4
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)


class ConvLayer(nn.Sequential):
    """[AvgPool2d(2)] -> EqualConv2d -> [LeakyReLU(0.2)]  (model.py:438-470; the bias exists only without activation)"""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=(1, 3, 3, 1), bias=True,
                 activate=True, stride=1, padding=1):
        layers = []
        if downsample:
            layers.append(nn.AvgPool2d(kernel_size=2, stride=2))
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(nn.LeakyReLU(0.2, inplace=True))
        super().__init__(*layers)


class ResBlock(nn.Module):
    """model.py:473-491"""

    def __init__(self, in_channel, out_channel, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3, stride=1, padding=1)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True, stride=1, padding=1)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False, stride=1, padding=0)

    def forward(self, input):
        out = self.conv1(input) * 1.4
        out = self.conv2(out) * 1.4
        skip = self.skip(input) * 1.4
        return (out + skip) / math.sqrt(2)


class Discriminator(nn.Module):
    """model.py:494-549: 1x1 stem (padding 1, so the map grows by 2), residual blocks down to 4x4 (+2), minibatch
    standard deviation over groups of 4, 3x3 conv, two linears -> one logit per image."""

    def __init__(self, size, ndf, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        channels = {4: ndf * 2, 8: ndf * 2, 16: ndf, 32: ndf, 64: ndf // 2, 128: ndf // 2}
        convs = [ConvLayer(3, channels[size], 1, stride=1, padding=1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3, stride=1, padding=1)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))

    def forward(self, input):
        out = self.convs(input) * 1.4
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        stddev = out.view(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2)
        stddev = stddev.repeat(group, 1, height, width)
        out = torch.cat([out, stddev], 1)
        out = self.final_conv(out) * 1.4
        return self.final_linear(out.view(batch, -1))


class uorfGanModel(uorfNoGanModel):
    """reference models/uorf_gan_model.py (forward :137-208, backward :240-245, forward_disc :247-311)."""

    def __init__(self, opt, precision: str = "fp16x3", with_perceptual: bool = True):
        super().__init__(opt, precision=precision, with_perceptual=with_perceptual)
        self.loss_names = ["recon", "perc", "gan", "d_real", "d_fake"]
        self.loss_recon = self.loss_perc = self.loss_gan = self.loss_d_real = self.loss_d_fake = 0
        self.weight_gan = 0
        self._render_only = False
        # init_type=None in the reference (:107): the discriminator keeps its own N(0,1) initialisation
        self.netDisc = Discriminator(size=opt.supervision_size, ndf=opt.ndf).to(self.device)
        toggle_grad(self.netDisc, False)          # the generator step never needs discriminator weight gradients
        if self.isTrain:
            self.d_optimizer = optim.Adam(self.netDisc.parameters(), lr=opt.d_lr, betas=(0.0, 0.9))
            self.d_scheduler = exp_decay_schedule_with_warmup(self.d_optimizer, opt.warmup_steps, opt.attn_decay_steps)

    # ---- generator step --------------------------------------------------------------------------------------
    def _weight_gan(self, epoch):
        opt = self.opt
        return opt.weight_gan if epoch >= opt.gan_train_epoch + opt.gan_in else 0

    def _graph_key_extra(self, epoch):
        return (self._weight_gan(epoch),)

    def forward(self, epoch=0):
        self.weight_gan = self._weight_gan(epoch)
        super().forward(epoch)

    def _extra_generator_loss(self, x_recon, x, epoch):
        """weight_gan * softplus(-D(x_recon)) (:187-188).  With weight_gan == 0 the reference still runs the discriminator
        and multiplies by zero; here the term (and the discriminator forward) is skipped."""
        if self._render_only or self.weight_gan == 0:
            self.loss_gan = 0
            return None
        self.loss_gan = self.weight_gan * g_nonsaturating_loss(self.netDisc(x_recon))
        return self.loss_gan

    def backward(self):
        super().backward()
        if self.weight_gan > 0:
            self.loss_gan = self.loss_gan.detach() / self.weight_gan

    # ---- discriminator step ----------------------------------------------------------------------------------
    def forward_disc(self, it=0):
        """One discriminator update on (x, render of the current scene); R1 penalty when it % 32 == 0 (:247-311)."""
        keep = (self.loss_recon, self.loss_perc, self.loss_gan, self.with_perceptual)
        self._render_only, self.with_perceptual = True, False             # the reference's forward_disc evaluates no losses
        try:
            with torch.no_grad():
                uorfNoGanModel.forward(self, epoch=self.opt.percept_in + 1)   # density noise 0, as in the reference (:291)
        finally:
            self._render_only = False
            self.loss_recon, self.loss_perc, self.loss_gan, self.with_perceptual = keep
        fake_img = self.x_recon
        real_img = self.x_sup.clone()
        toggle_grad(self.netDisc, True)
        fake_pred = self.netDisc(fake_img)
        real_pred = self.netDisc(real_img)
        d_loss_real, d_loss_fake = d_logistic_loss(real_pred, fake_pred)
        self.netDisc.zero_grad()
        (d_loss_real + d_loss_fake).backward()
        self.d_optimizer.step()
        if it % 32 == 0:
            real_img.requires_grad = True
            real_pred = self.netDisc(real_img)
            r1_loss = d_r1_loss(real_pred, real_img)
            self.netDisc.zero_grad()
            (self.opt.weight_r1 * r1_loss).backward()
            self.d_optimizer.step()
            self.loss_r1 = r1_loss.detach()
        toggle_grad(self.netDisc, False)
        self.loss_d_real, self.loss_d_fake = d_loss_real.detach(), d_loss_fake.detach()

    def get_current_losses(self):
        out = {}
        for n in self.loss_names:
            v = getattr(self, "loss_" + n)
            out[n] = float(v.detach()) if torch.is_tensor(v) else float(v)
        return out

    # ---- checkpoints: the reference stores only Encoder / SlotAttention / Decoder (model_names, :90); the discriminator
    # and its optimizer are saved next to them so an adversarial run can resume ----
    def save_networks(self, save_dir, suffix="latest"):
        super().save_networks(save_dir, suffix)
        torch.save({k: v.detach().cpu() for k, v in self.netDisc.state_dict().items()}, os.path.join(save_dir, f"{suffix}_net_Disc.pth"))
        if self.isTrain:
            torch.save(self.d_optimizer.state_dict(), os.path.join(save_dir, f"{suffix}_optimizer_disc.pth"))

    def load_networks(self, save_dir, suffix="latest", with_optimizer=True):
        super().load_networks(save_dir, suffix, with_optimizer)
        p = os.path.join(save_dir, f"{suffix}_net_Disc.pth")
        if os.path.exists(p):
            self.netDisc.load_state_dict(torch.load(p, map_location=str(self.device)))
        p = os.path.join(save_dir, f"{suffix}_optimizer_disc.pth")
        if with_optimizer and self.isTrain and os.path.exists(p):
            self.d_optimizer.load_state_dict(torch.load(p, map_location=str(self.device)))
