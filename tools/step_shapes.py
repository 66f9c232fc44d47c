"""Per-(op, shape) device time of one eagerly launched UNet step: CUDA events around every call of the functional
layer (kernels.py).  Usage: python tools/step_shapes.py [c2|c3] [out.txt]"""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from vidseg_diffusion_b200 import configs, kernels as K, linear
from vidseg_diffusion_b200.linear import Split

wlname = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[wlname]
cfg = {"sd21": configs.SD21_UNET, "svd": configs.SVD_UNET}[wl["cfg"]]
dev = torch.device("cuda", 0)
sd = bench.make_state_dict(cfg)
with torch.device("meta"):
    model = bench.model_class(cfg)(**cfg)
model = model.to_empty(device=dev); model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True); model.eval()
clip = [a.to(dev) for a in bench.make_clip(wl, cfg, 1)]
kw = dict(num_video_frames=wl["frames"], y=clip[3]) if bench.is_video(cfg) else {}
run = lambda: model(clip[0], timesteps=clip[1], context=clip[2], **kw)
run(); run()

rec = collections.defaultdict(list)


def shp(a):
    if isinstance(a, Split):
        return "S" + "x".join(map(str, a.hi.shape))
    if isinstance(a, torch.Tensor):
        return "T" + "x".join(map(str, a.shape))
    if isinstance(a, K.ChannelCat):
        return "Cat" + "x".join(map(str, a.shape))
    if isinstance(a, torch.nn.Module):
        w = getattr(a, "weight", None)
        return type(a).__name__ + ("" if w is None else "x".join(map(str, w.shape)))
    if isinstance(a, torch.nn.Parameter):
        return "P" + "x".join(map(str, a.shape))
    if a is None:
        return "-"
    return str(a)


def wrap(mod, name):
    orig = getattr(mod, name)

    def wrapped(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig(*a, **k); e1.record()
        key = (name,) + tuple(shp(x) for x in a) + tuple(f"{kk}={shp(v)}" for kk, v in sorted(k.items()) if v is not None and v is not False)
        rec[key].append((e0, e1))
        return out
    setattr(mod, name, wrapped)


# leaf ops only (linear -> gemm_split is wrapped at gemm_split; dense calls linear)
for n in ("gemm_split", "gemm_split_seg", "linear_geglu", "conv2d", "conv_temporal", "attention", "temporal_attention", "layer_norm_split",
          "geglu_split", "group_norm_split", "upsample_nearest2x_split", "image_split", "split"):
    wrap(K, n)

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
rows = []
for key, evs in rec.items():
    ms = sum(a.elapsed_time(b) for a, b in evs)
    rows.append((ms, key, len(evs)))
tot = sum(r[0] for r in rows)
lines = [f"{wlname}: {sum(r[2] for r in rows)} op calls, {tot:.2f} ms inside ops, {e0.elapsed_time(e1):.2f} ms for the eager step"]
byop = collections.defaultdict(float)
for ms, key, cnt in rows:
    byop[key[0]] += ms
lines += [f"  {k:28s} {v:8.3f} ms" for k, v in sorted(byop.items(), key=lambda t: -t[1])]
for ms, key, cnt in sorted(rows, reverse=True)[:70]:
    lines.append(f"{ms:8.3f} ms n={cnt:3d} avg={1e3 * ms / cnt:8.1f} us  {' '.join(key)}")
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
