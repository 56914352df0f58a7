"""Generate tests/golden/augment_chain.npz by running the UNMODIFIED reference's wrapper stack of the `augment` network
of BASELINE config 5 (mdir/examples/iccv23/parameters/finetune.yml:13):

    meanstd_post:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:[[0.485,0.456,0.406],[0.229,0.224,0.225]],clahepost:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:1.0,cir_ratio_pass_through:0.25:anc

through the reference's own `initialize_wrappers` / `Compose` on MetadataTensor inputs. The generator between the
wrappers is replaced by an exactly reproducible stand-in (flip + scale by 0.5: no conv round-off), so the fixture pins
the wrappers' arithmetic and routing bit for bit. Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_chain.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WRAPPERS = ("meanstd_post:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:[[0.485,0.456,0.406],[0.229,0.224,0.225]],"
            "clahepost:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:1.0,cir_ratio_pass_through:0.25:anc")


def stand_in_generator(images):
    return [None if x is None else x.flip(-1) * 0.5 for x in images]


def main():
    import torch
    ref_harness.load_reference()
    from mdir.components.data import wrapper as W
    from mdir.tools import tensors as T
    compose = W.initialize_wrappers(WRAPPERS, "cpu")
    rs = np.random.RandomState(51)
    names = ["img_%03d" % i for i in range(12)]
    labels = ["anc" if i % 3 == 0 else ("pos" if i % 3 == 1 else "neg") for i in range(12)]
    xs = [(rs.rand(1, 3, 40, 56).astype(np.float32) * 2.0 - 1.0) for _ in names]
    inputs = [T.MetadataTensor(torch.from_numpy(x.copy()), {"image_label": [lab], "name": [nm]})
              for x, lab, nm in zip(xs, labels, names)]
    with torch.no_grad():
        outs = compose(inputs, stand_in_generator)
    passed = [bool(compose.wrappers[2]._passthrough(nm)) and lab == "anc" for nm, lab in zip(names, labels)]
    assert any(passed) and not all(passed)
    pack = {"x": np.stack(xs), "y": np.stack([o.numpy() for o in outs]), "passed": np.array(passed),
            "labels": np.array(labels), "names": np.array(names), "wrappers": np.array(WRAPPERS)}
    # MeanStdPost alone on a [n,3,h,w] batch with awkward statistics
    post = W.MeanStdPost("[[0.5,0.4,0.3],[0.5,0.25,0.2]]", "[[0.485,0.456,0.406],[0.229,0.224,0.225]]", "cpu")
    xb = (rs.randn(3, 3, 17, 23) * 2).astype(np.float32)
    pack["ms_x"] = xb
    pack["ms_y"] = post.postprocess(torch.from_numpy(xb.copy()), None, None).numpy()
    np.savez_compressed(os.path.join(GOLD, "augment_chain.npz"), **pack)
    print("passed through the generator:", [n for n, p in zip(names, passed) if p])


if __name__ == "__main__":
    main()
