"""Times the photometric-loss launches alone (C ABI, no autograd glue) and the Adam launch; target for ncu."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from gaussianavatars_b200 import _native as N

dev = torch.device("cuda:0")
W, H = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
img = torch.rand(3, H, W, device=dev)
gt = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev)
grad = torch.empty_like(img)
scratch = torch.empty(N.PHOTOMETRIC_SCRATCH_HEAD + 3 * img.numel(), device=dev)
loss = torch.empty(3, device=dev)
a = N.PhotometricArgs()
a.abi_version = N.ABI_VERSION
a.channels, a.height, a.width, a.gt_is_u8, a.lambda_dssim = 3, H, W, 1, 0.2
a.image, a.gt, a.grad, a.loss, a.scratch = img.data_ptr(), gt.data_ptr(), grad.data_ptr(), loss.data_ptr(), scratch.data_ptr()
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(): N.check(N.lib().gab200_photometric_loss(C.byref(a), C.c_void_p(stream)), "loss")
for _ in range(int(os.environ.get("WARM", 3))): run()
torch.cuda.synchronize()
K = int(os.environ.get("ITERS", 20))
for mode in ("warm_l2", "flushed"):
    tot = 0.0
    for _ in range(K):
        if mode == "flushed": flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    n = img.numel()
    algo = n * (4 + 1 + 12) + n * (12 + 4 + 1 + 4)
    print(json.dumps({"kernel": "gab200_photometric_loss (stats + grad launches)", "W": W, "H": H, "l2": mode, "us": round(tot / K * 1e3, 1),
                      "algorithmic_bytes": algo, "achieved_gbs": round(algo / (tot / K) / 1e6, 1), "loss": [round(v, 6) for v in loss.tolist()]}), flush=True)

# ---- Adam launch alone (C ABI, one launch for six arrays of a 150k-splat model)
P = int(os.environ.get("P", 150000))
sizes = [P * 3, P * 3, P * 45, P, P * 3, P * 4]
bufs = [[torch.randn(s, device=dev) for _ in range(2)] + [torch.zeros(s, device=dev) for _ in range(2)] for s in sizes]
segs = (N.AdamSegment * len(sizes))(*[N.AdamSegment(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(), s, 1e-3)
                                       for b, s in zip(bufs, sizes)])
step = [0]
def adam():
    step[0] += 1
    N.check(N.lib().gab200_adam_step(len(sizes), segs, step[0], 0.9, 0.999, 1e-15, C.c_void_p(stream)), "adam")
for _ in range(3): adam()
torch.cuda.synchronize()
for mode in ("warm_l2", "flushed"):
    tot = 0.0
    for _ in range(K):
        if mode == "flushed": flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); adam(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    nbytes = sum(sizes) * 28
    print(json.dumps({"kernel": "gab200_adam_step (one launch, 6 arrays)", "elements": sum(sizes), "l2": mode, "us": round(tot / K * 1e3, 1),
                      "algorithmic_bytes": nbytes, "achieved_gbs": round(nbytes / (tot / K) / 1e6, 1)}), flush=True)
