"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). Used by
tools/gen_golden.py to produce tests/golden/* fixtures and by `-m "not gpu"` tests
that are skipped when the reference is absent. Recipe follows SURVEY.md App. B.
"""
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mdir"))


def load_reference():
    """Return the reference's `hubconf` module (CPU). Stubs four optional, absent modules."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    for m in ["h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
              "matplotlib.patches", "matplotlib.ticker", "imageio", "graphviz"]:
        if m not in sys.modules:
            sys.modules[m] = MagicMock(name=m)
    sys.modules["matplotlib"].rcParams = {"font.size": 10}
    if isinstance(sys.modules["h5py"], MagicMock):      # default_loader does isinstance(path, h5py.Dataset)
        sys.modules["h5py"].Dataset = type("Dataset", (), {})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import hubconf  # noqa: side effect: torch.set_num_threads(3) (mdir/stages/validate.py:10)
    return hubconf
