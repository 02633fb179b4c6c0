"""The reference README's quick start, timed end to end: get_mei_dict -> InvarianceManifold.learn(0) -> match([1, 2]) on demo1 at 100x100
(BASELINE configs[0] + configs[1]).

    python tools/quickstart.py ours                # laminr_b200 on cuda:0 (GPU box)
    python tools/quickstart.py reference [epochs]  # the unmodified reference on the host cores (build container only: needs /root/reference)

Prints one JSON line: wall clock of the three calls, epochs / steps taken, and summary statistics of the results (the pixel values of two runs that
took thousands of stochastic steps on different hardware are not comparable; the activations they reach are).
"""
import contextlib, io, json, os, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

which = sys.argv[1] if len(sys.argv) > 1 else "ours"
max_epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if which == "reference":
    from oracle.ref_harness import import_reference
    laminr = import_reference()
    device = "cpu"
    torch.set_num_threads(os.cpu_count() or 1)
else:
    import laminr_b200 as laminr
    device = "cuda"


def sync():
    if device == "cuda":
        torch.cuda.synchronize()


torch.manual_seed(0)
np.random.seed(0)
c = dict(pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
out = {"impl": which, "device": device, "threads": os.cpu_count() if device == "cpu" else None}
with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
    model = laminr.neuron_models.simulated("demo1", img_res=[100, 100]).to(device)
    sync(); t0 = time.perf_counter()
    meis = laminr.get_mei_dict(model, [1, 100, 100], **c)
    sync(); t1 = time.perf_counter()
    im = laminr.InvarianceManifold(model, meis, **c)
    imgs, acts = im.learn(0, num_max_epochs=max_epochs)
    sync(); t2 = time.perf_counter()
    aimgs, aacts = im.match([1, 2], num_epochs=max_epochs)
    sync(); t3 = time.perf_counter()
out.update({
    "get_mei_dict_s": t1 - t0, "learn_s": t2 - t1, "match_s": t3 - t2, "total_s": t3 - t0,
    "mei_activations": [float(meis[i]["activation"]) for i in range(3)],
    "mask_pixels_over_half": [int((np.asarray(meis[i]["mask"]) > 0.5).sum()) for i in range(3)],
    "learn_epochs": getattr(im, "learn_epochs_run", None), "learn_steps": getattr(im, "learn_steps_taken", None),
    "template_acts_mean_min": [float(np.mean(acts)), float(np.min(acts))],
    "match_epochs": getattr(im, "match_epochs_run", None),
    "aligned_acts_mean_per_target": [float(x) for x in np.asarray(aacts).mean(axis=1)],
    "shapes": [list(np.asarray(imgs).shape), list(np.asarray(aimgs).shape)],
})
print(json.dumps(out))
