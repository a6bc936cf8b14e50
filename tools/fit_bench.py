"""PSFNet-fitting workload (1_fit_psfnet.py: bs = 64 points x spp = 20000 rays per iteration, ks = 21): host + device time of
one PSFNet.get_training_data call through the reference-facing API, and where the host time goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdirt_b200 import lens_file
from sdirt_b200.deeplens import PSFNet

def main():
    dev = torch.device("cuda:0")
    lens = PSFNet(lens_file("rf50mm"), sensor_res=(512, 768), kernel_size=21, device=dev)
    for numerics in (None, "adaptive"):
        lens.numerics = numerics
        torch.manual_seed(0); np.random.seed(0)
        for _ in range(5):
            lens.get_training_data(bs=64, spp=20000)
        torch.cuda.synchronize()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            inp, psf = lens.get_training_data(bs=64, spp=20000)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"numerics={numerics}: get_training_data(bs=64, spp=20000): {dt * 1e3:.3f} ms / call = {64 * 20000 / dt:.3e} rays/s, "
              f"{64 / dt:.0f} PSFs/s")
    # whole fitting iterations (1_fit_psfnet.py / psfnet.py:101-167): ray-traced targets + MLP forward/backward/AdamW step
    import tempfile
    lens.numerics = "adaptive"
    with tempfile.TemporaryDirectory() as tmp:
        for use_graph in (False, True):
            lens.train_psfnet(iters=5, bs=64, spp=20000, evaluate_every=10 ** 9, result_dir=tmp, graph=use_graph)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_it = 200
            lens.train_psfnet(iters=n_it - 1, bs=64, spp=20000, evaluate_every=10 ** 9, result_dir=tmp, graph=use_graph)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n_it
            print(f"train_psfnet(bs=64, spp=20000, graph={use_graph}): {dt * 1e3:.3f} ms / iteration = {1 / dt:.0f} it/s "
                  f"(the reference's 90 k-iteration fit: {90000 * dt / 60:.1f} min)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        lens.get_training_data(bs=64, spp=20000)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    t0 = time.perf_counter()
    inp, psf = lens.get_test_data(bs=1024, spp=65536)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"get_test_data(bs=1024, spp=65536): {dt * 1e3:.1f} ms = {1024 * 65536 / dt:.3e} rays/s")

main()
