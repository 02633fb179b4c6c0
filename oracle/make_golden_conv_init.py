"""TEST INFRASTRUCTURE ONLY -- what `torch.manual_seed(0)` + the reference's vendored neuralpredictors constructors produce for the conv-core response
model of BASELINE config 4 (`Stacked2dCore(1, 32, 9, 7, layers=3, stack=-1, pad_input=True)`, `FullGaussian2d((32, 200, 200), 128,
bias=True)`; cores/conv2d.py:40-68, readouts/gaussian.py:452-469) and for a small variant: the sha256 of every state_dict entry and of the CPU
generator state after construction.  The host mirror `laminr_b200/neuron_models/conv_models.py` must reproduce both (same draws, same order).

    python -m oracle.make_golden_conv_init        ->  tests/golden/conv_init.npz
"""
import hashlib
import os
import warnings

import numpy as np
import torch

from oracle.make_golden import OUT
from oracle.ref_harness import import_reference

CONFIGS = {
    "config4": (dict(input_channels=1, hidden_channels=32, input_kern=9, hidden_kern=7, layers=3, stack=-1, pad_input=True),
                dict(in_shape=(32, 200, 200), outdims=128, bias=True)),
    "small": (dict(input_channels=2, hidden_channels=8, input_kern=5, hidden_kern=3, layers=2, stack=-1, pad_input=True),
              dict(in_shape=(8, 24, 20), outdims=6, bias=True)),
}


def digest(core_cls, readout_cls):
    """{config}_{core|readout}.{state_dict key} -> sha256 (uint8[32]); {config}_rng_after -> sha256 of the generator state."""
    out = {}
    for tag, (ckw, rkw) in CONFIGS.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.manual_seed(0)
            core = core_cls(**ckw)
            readout = readout_cls(**rkw)
        for part, mod in (("core", core), ("readout", readout)):
            for k, v in mod.state_dict().items():
                out[f"{tag}_{part}.{k}"] = np.frombuffer(hashlib.sha256(v.detach().cpu().contiguous().numpy().tobytes()).digest(), np.uint8)
        out[f"{tag}_rng_after"] = np.frombuffer(hashlib.sha256(torch.get_rng_state().numpy().tobytes()).digest(), np.uint8)
    return out


def main():
    import_reference()
    from neuralpredictors.layers.cores.conv2d import Stacked2dCore
    from neuralpredictors.layers.readouts.gaussian import FullGaussian2d

    out = digest(Stacked2dCore, FullGaussian2d)
    np.savez_compressed(os.path.join(OUT, "conv_init.npz"), **out)
    print(len(out), "entries:", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
