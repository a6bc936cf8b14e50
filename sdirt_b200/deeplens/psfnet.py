"""Host-side mirror of the reference's `PSFNet` (deeplens/psfnet.py): the PSF-bank workload generators
(`get_training_data`, `get_test_data`), the fitting loop, the MLP prediction `pred`, and the spatially varying dual-pixel
`render`.  Ray tracing and rendering run on libsdirt_engine; inside `render` the MLP is the engine's fused tensor-core kernel
(sdirt_mlp_fused_pred: one launch per band of pixels) with the library-GEMM chain as the route for shapes that kernel is not
compiled for; `pred` on its own and the fitting step are plain torch."""
import logging
import os

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

    fit_graph = True                 # capture forward + loss + backward of the fitting step in a CUDA graph
    fit_device_data = False          # train_psfnet: draw the batch's (x, y, z) and the shared pupil samples on the DEVICE and keep
                                     # the evaluation bank (get_test_data) resident between evaluations -- no host RNG, no
                                     # per-iteration upload.  Off by default: a seeded run then draws what the reference draws.

    def _training_data_device(self, bs, spp):
        """get_training_data with every draw on the device generator (same distributions, psfnet.py:170-199)."""
        foc_z = float(np.random.choice(self.foc_z_arr))
        x = (torch.rand(bs, device=self.device) - 0.5) * 2
        y = (torch.rand(bs, device=self.device) - 0.5) * 2
        zg = torch.clamp(torch.randn(bs, device=self.device), min=-3, max=3)
        z = torch.where(zg > 0, (1 - foc_z) * zg / 3 + foc_z, torch.where(zg < 0, foc_z * zg / 3 + foc_z, torch.zeros_like(zg)))
        inp = torch.stack((x, y, z), dim=-1)
        points = torch.stack((x, y, self.z2depth(z)), dim=-1)
        prev = getattr(self, "sample_rng", "cpu")
        self.sample_rng = "cuda"
        try:
            return inp, self.psf(points=points, ks=self.kernel_size, spp=spp)
        finally:
            self.sample_rng = prev

    def train_psfnet(self, iters=10000, bs=128, lr=1e-4, spp=2048, evaluate_every=1000, result_dir="./results/temp", graph=None):
        """Fit the PSF MLP to ray-traced PSFs generated on the fly (psfnet.py:101-167); no plotting.

        The targets come from the engine and stay on the device; the loss history is read back once at the end.  With
        `graph` (default `fit_graph`) the forward, loss and backward of the step are captured once in a CUDA graph and
        replayed (about 70 kernel launches per iteration become one); GradScaler.step / update and the scheduler run
        eagerly, as torch's AMP requires."""
        psfnet = self.psfnet
        psfnet.train()
        l2 = nn.MSELoss(reduction="mean")
        use_graph = self.fit_graph if graph is None else bool(graph)
        # same AdamW update either way; the fused implementation takes GradScaler's inf check on the device (no host sync)
        optim = torch.optim.AdamW(psfnet.parameters(), lr, fused=True) if use_graph else torch.optim.AdamW(psfnet.parameters(), lr)
        sche = torch.optim.lr_scheduler.CosineAnnealingLR(optim, T_max=max(int(iters) // 3, 1), eta_min=0)
        scaler = torch.amp.GradScaler("cuda")
        ks = self.kernel_size
        losses = []
        self.eval_history = []
        g = None
        if use_graph:
            static_inp = torch.zeros((bs, 3), device=self.device)
            static_psf = torch.zeros((bs, ks, ks), device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                               # warm-up off the capture (no optimizer step: the fit
                for _ in range(2):                                       # itself starts from the untouched weights)
                    with torch.autocast(device_type="cuda"):
                        warm = l2(psfnet(static_inp), static_psf)
                    scaler.scale(warm).backward()
                    optim.zero_grad(set_to_none=True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                with torch.autocast(device_type="cuda"):
                    static_loss = l2(psfnet(static_inp), static_psf)
                scaler.scale(static_loss).backward()
        for i in range(iters + 1):
            inp, psf = self._training_data_device(bs, spp) if self.fit_device_data else self.get_training_data(bs=bs, spp=spp)
            inp, psf = inp.to(self.device), psf.to(self.device)
            if g is not None:
                static_inp.copy_(inp)
                static_psf.copy_(psf)
                g.replay()                                               # gradients are rewritten, not accumulated
                losses.append(static_loss.detach().clone())
            else:
                with torch.autocast(device_type="cuda"):
                    loss = l2(psfnet(inp), psf)
                optim.zero_grad()
                scaler.scale(loss).backward()
                losses.append(loss.detach())
            scaler.step(optim)
            scaler.update()
            sche.step()
            if (i + 1) % evaluate_every == 0:
                # psfnet.py:135-166 without the PNG: checkpoint, then L1 / L2 of the sum-normalised prediction on get_test_data
                # (1024 grid points x 65536 rays, traced by the engine), logged in the reference's format
                os.makedirs(result_dir, exist_ok=True)
                torch.save(psfnet.state_dict(), f"{result_dir}/iter{i + 1}_PSFNet_{self.model_name}.pkl")
                with torch.no_grad(), torch.autocast(device_type="cuda"):
                    psfnet.eval()
                    if self.fit_device_data and getattr(self, "_test_bank", None) is not None and self._test_bank[0] == float(self.d_sensor):
                        tin, tpsf = self._test_bank[1], self._test_bank[2]          # the evaluation bank, traced once
                    else:
                        tin, tpsf = self.get_test_data()
                        tin, tpsf = tin.to(self.device), tpsf.to(self.device)
                        if self.fit_device_data:
                            self._test_bank = (float(self.d_sensor), tin, tpsf)
                    tpred = psfnet(tin)
                    tpsf = tpsf / tpsf.sum(-1).sum(-1).unsqueeze(-1).unsqueeze(-1)
                    tpred = tpred / tpred.sum(-1).sum(-1).unsqueeze(-1).unsqueeze(-1)
                    l1_loss, l2_loss = nn.L1Loss(reduction="mean")(tpred, tpsf), l2(tpred, tpsf)
                    logging.info(f"{i}, {l1_loss.item()}, {l2_loss.item()}")
                    self.eval_history.append((i, l1_loss.item(), l2_loss.item()))
                    psfnet.train()
        os.makedirs(result_dir, exist_ok=True)
        torch.save(psfnet.state_dict(), f"{result_dir}/PSFNet_{self.model_name}.pkl")          # psfnet.py:167: what the configs load
        return torch.stack(losses).float().cpu().tolist()

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

    # ---- banded render: the per-pixel PSF tensor only ever exists for one band of rows -------------------
    render_band_rows = 16            # rows per band (the render kernels' tile height)
    render_band_pixels = 98304       # pixels per band batch: images are grouped until a band holds about this many
    render_band_pixels_fused = 393216  # budget of the fused engine's automatic band shape (2 x 0.7 GB of fp16 PSFs in flight)
    render_overlap = True            # pack + convolve band i on a second stream while the MLP of band i + 1 runs
    mlp_engine = "fused"             # "fused": the whole of `pred` for a band as one tcgen05 kernel (csrc/mlp_fused.cuh);
                                     # "cublas": input-layer kernel + torch GEMM chain + pack kernel (same results bit for bit)

    def _fused_band_shape(self, N, H, W):
        """(rows, images) of a band for the fused engine.  Its persistent kernel gives every CTA pair groups of 128 pixels, so a
        band costs ceil(pixels / 128 / (SMs / 2)) rounds: pick the shape under the pixel budget that wastes the least of its
        last round (98304 pixels = 10.4 rounds cost 11; 196608 = 20.8 cost 21) -- weighed against what a short band costs the
        render kernel, which walks 32-pixel strips downwards and pays ~3 rows' time (KS image rows to fetch, one copy latency)
        wherever a strip of the band ends.  The convolution is ~5 % of a band's time."""
        if "render_band_rows" in self.__dict__ or "render_band_pixels" in self.__dict__:      # set by hand: keep
            rows = max(1, min(int(self.render_band_rows), H))
            return rows, max(1, min(N, int(self.render_band_pixels) // (rows * W)))
        sms = E.lib().sdirt_device_sm_count()
        slots = max(1, (sms if sms > 0 else 148) // 2)
        best = None
        for rows in sorted({min(H, r) for r in range(16, 129, 16)}):
            for nb in range(1, N + 1):
                px = nb * rows * W
                if px > max(int(self.render_band_pixels_fused), rows * W) and nb > 1:
                    break
                rounds = -(-px // 128) / slots
                mlp_eff, conv_eff = rounds / -(-rounds // 1), rows / (rows + 3.0)
                score = (round(1.0 / (0.95 / mlp_eff + 0.05 / conv_eff), 3), px)
                if best is None or score > best[0]:
                    best = (score, rows, nb)
        return best[1], best[2]

    def _mlp_fused(self):
        """The MLP packed for sdirt_mlp_fused_pred, cached on the parameters' versions."""
        lin = [m for m in self.psfnet.net if isinstance(m, nn.Linear)]
        key = tuple((m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for m in lin)
        cache = getattr(self, "_mlp_fused_cache", None)
        if cache is None or cache[0] != key:
            # The fused kernel is compiled for kernel sizes 7 / 11 / 21 and layer widths that tile its MMA shape; any other
            # PSFNet (ks = 9, 31, 35, 51 ...) takes the library-GEMM route (generic pack + generic render kernels): None here.
            try:
                packed = E.FusedMlp([(m.weight, m.bias) for m in lin]) if self.kernel_size in (7, 11, 21) else None
            except RuntimeError:
                packed = None
            cache = self._mlp_fused_cache = (key, packed)
        return cache[1]

    def _render_post_stream(self, device):
        st = getattr(self, "_post_stream", None)
        if st is None or st.device != torch.device(device):
            st = self._post_stream = torch.cuda.Stream(device=device)
        return st

    def _mlp_half_layers(self):
        """fp16 copies of the MLP's Linear layers, as CUDA autocast casts them on every call (psfnet_arch.py:52); cached on
        the parameters' versions.  The last layer's N is padded to a multiple of 8 (16-byte rows for cuBLAS)."""
        lin = [m for m in self.psfnet.net if isinstance(m, nn.Linear)]
        key = tuple((m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for m in lin)
        cache = getattr(self, "_mlp_half_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        first = (lin[0].weight.detach().half().contiguous(), lin[0].bias.detach().half().contiguous())
        chain = []
        for i, m in enumerate(lin[1:]):
            w, b = m.weight.detach().half(), m.bias.detach().half()
            if i == len(lin) - 2 and w.shape[0] % 8:
                padn = 8 - w.shape[0] % 8
                w = torch.cat((w, w.new_zeros(padn, w.shape[1])), 0)
                b = torch.cat((b, b.new_zeros(padn)), 0)
            chain.append((w.contiguous().t(), b.contiguous()))
        layers = (first, chain)
        self._mlp_half_cache = (key, layers)
        return layers

    @torch.no_grad()
    def _render_banded(self, img, depth, tone):
        """degamma (tone & 1) -> per-pixel PSFs -> gather-convolution -> gamma + clip (tone & 2), band by band; [N, 2C, H, W]."""
        if img.dim() != 4:
            raise NotImplementedError("PSFNet.render expects a batched [N, C, H, W] image")
        if not img.is_cuda:
            raise RuntimeError("PSFNet.render: the engine has no CPU path (img must be a CUDA tensor)")
        depth = depth + self.d_sensor                                     # the reference's d_sensor fix, :658
        N, C, H, W = img.shape
        ks = self.kernel_size
        z = self.depth2z(depth).reshape(N, H, W).float().contiguous()
        xs = torch.linspace(-1, 1, W).to(img.device)                      # made on the host, as the reference does (:684-688)
        ys = torch.linspace(1, -1, H).to(img.device)
        (w1, b1), chain = self._mlp_half_layers()
        img32 = img.float().contiguous()
        # the image as the strip-walking render kernel streams it (padded, degamma'd, fp16), once per call and not per band;
        # None for shapes that kernel does not take (other channel counts / widths / kernel sizes): then degamma once per call
        rec = E.render_pack_image(img32, ks, tone & 1)
        if rec is not None:
            tone &= ~1
        elif tone & 1:
            img32, tone = E.tone_degamma(img32, out=img32 if img32.data_ptr() != img.data_ptr() else None), tone & ~1
        rl, rr = torch.empty_like(img32), torch.empty_like(img32)
        use_fused = self.mlp_engine == "fused" and self._mlp_fused() is not None
        if use_fused:
            rows, nb = self._fused_band_shape(N, H, W)
        else:
            rows = max(1, min(int(self.render_band_rows), H))
            nb = max(1, min(N, int(self.render_band_pixels) // (rows * W)))
        # Two streams: the GEMM chain of band i + 1 (tensor cores) runs while band i is packed and convolved (memory pipes).
        # raw / psf buffers are double-buffered by hand so that no tensor crosses streams through the caching allocator.
        main = torch.cuda.current_stream(img.device)
        post = self._render_post_stream(img.device) if self.render_overlap else main
        n_out = chain[-1][0].shape[1]
        max_px = nb * rows * W
        fused = self._mlp_fused() if (use_fused and max_px % 4 == 0 and (rows * W) % 4 == 0) else None
        raw = [torch.empty((2 * max_px, n_out), device=img.device, dtype=torch.float16) for _ in range(2)] if fused is None else None
        psf = [torch.empty((max_px, 2, ks, ks), device=img.device, dtype=torch.float16) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        buf_free = [None, None]
        post.wait_stream(main)                                             # img32 / z / outputs are ready for the post stream
        i = 0
        for b0 in range(0, N, nb):
            nbb = min(nb, N - b0)
            for y0 in range(0, H, rows):
                nr = min(rows, H - y0)
                px = nbb * nr * W
                k = i & 1
                if fused is not None and px % 4 == 0:
                    if buf_free[k] is not None:
                        main.wait_event(buf_free[k])                      # band i - 2 has been convolved out of this buffer
                    fused.pred(xs, ys, z, b0, nbb, y0, nr, ks, out=psf[k][:px])
                    packed = True
                else:
                    if raw is None:
                        raw = [torch.empty((2 * max_px, n_out), device=img.device, dtype=torch.float16) for _ in range(2)]
                    h = E.mlp_input_layer(xs, ys, z, b0, nbb, y0, nr, w1, b1)
                    for wt, b in chain[:-1]:
                        h = torch._addmm_activation(b, h, wt)             # relu(h @ W^T + b), fp16 in / fp32 accumulate / fp16 out
                    if buf_free[k] is not None:
                        main.wait_event(buf_free[k])
                    torch._addmm_activation(chain[-1][1], h, chain[-1][0], out=raw[k][:2 * px])
                    packed = False
                ready[k].record(main)
                with torch.cuda.stream(post):
                    post.wait_event(ready[k])
                    if not packed:
                        E.psf_pack(raw[k][:2 * px], ks, out=psf[k][:px])
                    if rec is not None:
                        nrec = rec.numel() // N
                        E.render_local_psf_rows_packed(rec[b0 * nrec:(b0 + nbb) * nrec], (nbb, C, H, W), psf[k][:px].view(nbb, nr, W, 2, ks, ks),
                                                       ks, y0, rl[b0:b0 + nbb], rr[b0:b0 + nbb], tone=tone)
                    else:
                        E.render_local_psf_rows(img32[b0:b0 + nbb], psf[k][:px].view(nbb, nr, W, 2, ks, ks), ks, y0,
                                                rl[b0:b0 + nbb], rr[b0:b0 + nbb], tone=tone)
                    if post is not main:
                        buf_free[k] = torch.cuda.Event()
                        buf_free[k].record(post)
                i += 1
        main.wait_stream(post)
        return torch.cat((rl, rr), dim=1)

    @torch.no_grad()
    def render(self, img, depth, foc_dist, train=False):
        """[N, 6, H, W] dual-pixel image (left RGB, right RGB) from an all-in-focus image [N, 3, H, W] and a depth
        map [N, 1, H, W] in negative millimetres (psfnet.py:645-714).

        The reference builds the per-pixel PSF tensor [N,H,W,2,ks,ks] of the whole batch (`pred`) and then convolves.
        Here the image is walked in bands of rows.  `mlp_engine = "fused"` (default): ONE tcgen05 kernel per band takes the
        coordinates through all eleven layers to the flipped / stacked / normalised fp16 kernels (sdirt_mlp_fused_pred), and one
        kernel convolves them (degamma + gather-convolution + gamma + clip).  `mlp_engine = "cublas"`, and any PSFNet the fused
        kernel is not compiled for: first Linear (engine kernel), the 512-wide GEMM chain (cuBLAS, bias + ReLU in the epilogue),
        flip / stack / normalise (engine kernel), then the same convolution -- bit-identical kernels either way.  The PSFs of a
        band stay in the L2 between their producer and their consumer, and memory use does not grow with the image."""
        render = self._render_banded(img, depth, 1 if train else 3)
        if train:                                                          # noise sits between gamma and clip (:708-713)
            # one draw per call, in the reference's order: noise_range, randn_like, range1, range2 (psfnet.py:629-642)
            N, W = img.shape[0], img.shape[-1]
            noise_range = 0.05 * np.random.rand()
            randn = torch.randn_like(render)
            range1, range2 = (np.random.rand() / 2), (np.random.rand() / 2 + 0.5)
            weight = torch.linspace(range1, range2, W).to(img.device).repeat(N, 1)
            nr_t = torch.full((N,), noise_range, dtype=torch.float32, device=img.device)
            E.gamma_noise_clip(render, randn, nr_t, weight)
        return render

    @torch.no_grad()
    def render_focal_stack(self, aif, depth, foc_dists, train=True):
        """The data-generation loop of 2_dfdp_net.py:161-173 (`for i in range(bs): render(aif[i:i+1], ...)`, then cat) as one
        banded pass over the whole batch.  Random draws are made per image in the loop's order (noise_range, randn of one
        image, range1, range2), so a seeded run produces the images the reference's loop would."""
        N, C, H, W = aif.shape
        if not train:
            return self._render_banded(aif, depth, 3)
        lin = self._render_banded(aif, depth, 1)
        randn = torch.empty_like(lin)
        nr, ramps = [], []
        for i in range(N):
            nr.append(0.05 * np.random.rand())
            randn[i:i + 1].normal_()
            range1, range2 = (np.random.rand() / 2), (np.random.rand() / 2 + 0.5)
            ramps.append(torch.linspace(range1, range2, W))
        E.gamma_noise_clip(lin, randn, torch.tensor(nr, dtype=torch.float32).to(aif.device), torch.stack(ramps).to(aif.device))
        return lin

    @torch.no_grad()
    def render_via_pred(self, img, depth, foc_dist, train=False):
        """The reference's own order of operations (psfnet.py:681-713): `pred` for every pixel of the batch, then one
        convolution over the whole tensor.  Kept as the memory-hungry comparison path for tests and benchmarks."""
        depth = depth + self.d_sensor
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
        if train:
            render = self.gamma(render)
            render = self.noise(render, img.shape)
            render = torch.clip(render, 0.0, 1.0)
        return render

    @torch.no_grad()
    def compare_psf(self):
        """The data of PSFNet.compare_psf (psfnet.py:529-566) without its PNGs: for the three field points (0, 0), (0.4, 0.4),
        (0.8, 0.8) at 0.5 m and 20 m, the ray-traced (left, right) PSFs -- the right one from the mirrored point, flipped, as the
        reference obtains it -- and the network's prediction.  Returns {distance: (traced [3, 2, ks, ks], predicted [3, 2, ks, ks])}."""
        from .basics import GEO_SPP
        x = torch.Tensor([0, 0.4, 0.8])
        y = torch.Tensor([0, 0.4, 0.8])
        out = {}
        for d_ori in (-500.0, -20000.0):
            depth = d_ori + self.d_sensor
            inp = torch.stack((x, y, torch.full_like(x, depth)), dim=-1)
            psfl = self.psf(points=inp, ks=self.kernel_size, center=True, spp=GEO_SPP * 100).cpu()
            inp[..., 0] = inp[..., 0] * (-1)
            psfr = torch.flip(self.psf(points=inp, ks=self.kernel_size, center=True, spp=GEO_SPP * 100).cpu(), dims=[-1])
            z = float(self.depth2z(torch.tensor(depth)))
            net_in = torch.stack((x, y, torch.full_like(x, z)), dim=-1).repeat(1, 1, 1, 1).to(self.device)
            pred = self.pred(net_in).detach().float().cpu()[0, 0]                                   # [3, 2, ks, ks]
            out[int(d_ori)] = (torch.stack((psfl, psfr), dim=1), pred)
        return out

    def depth2z(self, depth):
        return torch.clamp((depth - self.d_min) / (self.d_max - self.d_min), min=0, max=1)

    def z2depth(self, z):
        return z * (self.d_max - self.d_min) + self.d_min
