"""Host-side mirror of the reference's `PSFNet` (deeplens/psfnet.py): the PSF-bank workload generators
(`get_training_data`, `get_test_data`), the MLP prediction `pred`, and the spatially varying dual-pixel `render`.
Ray tracing and rendering run on libsdirt_engine; the MLP is plain torch (cuBLAS)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _engine as E
from .optics import Lensgroup
from .psfnet_arch import MLP, initialize_weights

DMIN = 200      # [mm]
DMAX = 20000    # [mm]


class PSFNet(Lensgroup):
    def __init__(self, filename, model_name="mlp", kernel_size=11, sensor_res=(512, 512), device="cuda"):
        super().__init__(filename=filename, sensor_res=sensor_res, device=device)
        self.in_features = 4
        self.kernel_size = kernel_size
        self.model_name = model_name
        self.init_net()
        self.spp = 4096
        self.patch_size = 64
        self.psf_grid = [sensor_res[0] // self.patch_size, sensor_res[1] // self.patch_size]
        self.d_max = -DMAX
        self.d_min = -DMIN
        # psfnet.py:42-48: the sensor position is overridden per lens WITHOUT recomputing hfov
        if filename.find("rf35mm") != -1:
            self.d_sensor = 80.447
        elif filename.find("rf50mm") != -1:
            self.d_sensor = 62.25
        else:
            raise ValueError("filename is not correct: PSFNet hard-codes d_sensor for rf35mm / rf50mm only")
        self.foc_d_arr = np.array([-999.9, -1000, -1000.1], dtype=np.float32) + self.d_sensor
        self.foc_z_arr = (self.foc_d_arr - self.d_min) / (self.d_max - self.d_min)
        self.foc_d = np.array([-1000.0], dtype=np.float32) + self.d_sensor
        self.psf_shot_modeling = None

    # ---- network ---------------------------------------------------------------------------------
    def init_net(self):
        ks = self.kernel_size
        if self.model_name == "mlp":
            self.psfnet = MLP(in_features=3, out_features=ks ** 2, hidden_features=512, hidden_layers=8)
        else:
            raise Exception("Unsupported PSF network architecture.")     # mlpconv / siren: outside the hot path
        self.psfnet.apply(initialize_weights)
        self.psfnet.to(self.device)

    def load_net(self, net_path):
        net_dict = self.psfnet.state_dict()
        pretrain = torch.load(net_path, map_location=self.device)
        net_dict.update({k: v for k, v in pretrain.items() if k in net_dict and net_dict[k].shape == v.shape})
        self.psfnet.load_state_dict(net_dict)

    # ---- PSF-bank workload generators (psfnet.py:170-241) -------------------------------------------
    def _warp_z(self, z_gauss, foc_z):
        z = torch.zeros_like(z_gauss)
        z[z_gauss > 0] = (1 - foc_z) * z_gauss[z_gauss > 0] / 3 + foc_z
        z[z_gauss < 0] = foc_z * z_gauss[z_gauss < 0] / 3 + foc_z
        return z

    def get_training_data(self, bs=256, spp=4096):
        foc_z = np.random.choice(self.foc_z_arr)
        x = (torch.rand(bs) - 0.5) * 2
        y = (torch.rand(bs) - 0.5) * 2
        z = self._warp_z(torch.clamp(torch.randn(bs), min=-3, max=3), foc_z)
        inp = torch.stack((x, y, z), dim=-1)
        points = torch.stack((x, y, self.z2depth(z)), dim=-1)
        return inp, self.psf(points=points, ks=self.kernel_size, spp=spp)

    def get_test_data(self, bs=1024, spp=65536):
        foc_z = self.foc_z_arr[1]
        g = 32
        x, y = torch.meshgrid(torch.linspace(-1 + 1 / (2 * g), 1 - 1 / (2 * g), g),
                              torch.linspace(1 - 1 / (2 * g), -1 + 1 / (2 * g), g), indexing="xy")
        x, y = x.reshape(-1), y.reshape(-1)
        z = self._warp_z(torch.linspace(-3, 3, bs), foc_z)
        inp = torch.stack((x, y, z), dim=-1)
        points = torch.stack((x, y, self.z2depth(z)), dim=-1)
        return inp, self.psf(points=points, ks=self.kernel_size, spp=spp)

    def train_psfnet(self, iters=10000, bs=128, lr=1e-4, spp=2048, evaluate_every=1000, result_dir="./results/temp"):
        """Fit the PSF MLP to ray-traced PSFs generated on the fly (psfnet.py:101-167); no plotting."""
        psfnet = self.psfnet
        psfnet.train()
        l2 = nn.MSELoss(reduction="mean")
        optim = torch.optim.AdamW(psfnet.parameters(), lr)
        sche = torch.optim.lr_scheduler.CosineAnnealingLR(optim, T_max=max(int(iters) // 3, 1), eta_min=0)
        scaler = torch.amp.GradScaler("cuda")
        loss_hist = []
        for i in range(iters + 1):
            inp, psf = self.get_training_data(bs=bs, spp=spp)
            inp, psf = inp.to(self.device), psf.to(self.device)
            with torch.autocast(device_type="cuda"):
                loss = l2(psfnet(inp), psf)
            optim.zero_grad()
            scaler.scale(loss).backward()
            scaler.step(optim)
            scaler.update()
            sche.step()
            loss_hist.append(loss.item())
            if (i + 1) % evaluate_every == 0:
                torch.save(psfnet.state_dict(), f"{result_dir}/iter{i + 1}_PSFNet_{self.model_name}.pkl")
        return loss_hist

    # ---- prediction and rendering (psfnet.py:317-336, 589-726) --------------------------------------
    def pred(self, inp):
        """Per-pixel L/R PSFs from the MLP: R is the mirrored evaluation, flipped.  Mutates inp[..., 0] like the
        reference does (psfnet.py:328)."""
        psfl = self.psfnet(inp)
        inp[..., 0] = inp[..., 0] * (-1)
        psfr = torch.flip(self.psfnet(inp), dims=[-1])
        psf = torch.stack((psfl, psfr), dim=-3)
        psf = psf / (psf.sum(-1).sum(-1).unsqueeze(-1).unsqueeze(-1) + 1e-9)
        assert psf.shape[-1] == self.kernel_size
        return psf

    def fit_degamma(self, x):
        a1, b1, c1 = 0.89129432, 0.27217316, -0.00246187
        a2, b2, c2 = 5.94018909e-01, 1.20060450e+01, -5.24983855e-03
        l1 = 1 / (1 / (a1 * x + b1) + c1)
        l2 = 1 / (1 / (a2 * x + b2) + c2)
        ratio_x = x / 100
        ratio_x[ratio_x > 1] = 1
        return l2 * ratio_x + l1 * (1 - ratio_x)

    def degamma(self, img_gamma):
        return self.fit_degamma(img_gamma * 255.)

    def fit_gamma(self, l):
        a1, b1, c1 = 0.89129432, 0.27217316, -0.00246187
        a2, b2, c2 = 5.94018909e-01, 1.20060450e+01, -5.24983855e-03
        x1 = (1 / (1 / (l + 1e-9) - c1) - b1) / a1
        x2 = (1 / (1 / (l + 1e-9) - c2) - b2) / a2
        ratio_x = ((x1 + x2) / 2) / 100
        ratio_x[ratio_x > 1] = 1
        return x2 * ratio_x + x1 * (1 - ratio_x)

    def gamma(self, img_degamma):
        return self.fit_gamma(img_degamma) / 255.

    def noise(self, render, shape):
        N, C, H, W = shape
        noise_range = 0.05 * np.random.rand()
        noise_map = torch.randn_like(render) * noise_range
        range1, range2 = (np.random.rand() / 2), (np.random.rand() / 2 + 0.5)
        weight_l = torch.linspace(range1, range2, W).repeat(N, C, H, 1)
        weight_map = torch.cat([weight_l, torch.flip(weight_l, [-1])], dim=1).to(render.device)
        render += noise_map * weight_map
        return render

    @torch.no_grad()
    def render(self, img, depth, foc_dist, train=False):
        """[N, 6, H, W] dual-pixel image (left RGB, right RGB) from an all-in-focus image [N, 3, H, W] and a depth
        map [N, 1, H, W] in negative millimetres (psfnet.py:645-714).  degamma, the per-pixel gather-convolution,
        gamma and the final clip run in ONE engine kernel."""
        if img.dim() != 4:
            raise NotImplementedError("PSFNet.render expects a batched [N, C, H, W] image")
        depth = depth + self.d_sensor                                     # the reference's d_sensor fix, :658
        N, C, H, W = img.shape
        z = self.depth2z(depth).squeeze(1)
        x, y = torch.meshgrid(torch.linspace(-1, 1, W), torch.linspace(1, -1, H), indexing="xy")
        x, y = x.unsqueeze(0).repeat(N, 1, 1).to(img.device), y.unsqueeze(0).repeat(N, 1, 1).to(img.device)
        o = torch.stack((x, y, z), -1).float()
        psf = self.pred(o)
        if psf.dtype not in (torch.float16, torch.float32):
            psf = psf.float()
        rl, rr = E.render_local_psf(img.float().contiguous(), psf.contiguous(), self.kernel_size, tone=1 if train else 3)
        render = torch.cat((rl, rr), dim=1)
        if train:                                                          # noise sits between gamma and clip
            render = self.gamma(render)
            render = self.noise(render, img.shape)
            render = torch.clip(render, 0.0, 1.0)
        return render

    def depth2z(self, depth):
        return torch.clamp((depth - self.d_min) / (self.d_max - self.d_min), min=0, max=1)

    def z2depth(self, z):
        return z * (self.d_max - self.d_min) + self.d_min
