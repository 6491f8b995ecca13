"""Import shim for the UNMODIFIED reference installed under baseline/_ref (git-ignored; it travels to the GPU box):

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/refcopy            # /tmp/refcopy = copy of /root/reference (read-only source tree)
    cp /root/reference/pretrained/MonkeyHand/*-small.pth baseline/_ref/pretrained/MonkeyHand/

Used only by baseline legs (bench.py `torch_gpu_baseline`) and by tests that exercise the drop-in seam on the
reference's real classes (tests/test_gpu_reference_dropin.py).  Nothing of the product path imports it.

The reference's *unused* imports pull in packages that are not in this image (matplotlib, imgaug, streamlit, yacs,
ruamel.yaml, seaborn, inquirer, streamlit_option_menu — SURVEY.md section 8c): empty stub modules stand in for them;
none is called on the inference path.  On a GPU box the reference then runs its native CUDA path unchanged."""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
WEIGHTS = os.path.join(REF, "pretrained", "MonkeyHand")


class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()


def _stub_getattr(n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Any


def available():
    return os.path.isdir(os.path.join(REF, "jarvis", "hybridnet"))


def import_reference():
    """Make `import jarvis...` resolve to baseline/_ref.  Returns False when the install is absent."""
    if not available():
        return False
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "imgaug",
                 "imgaug.augmenters", "imgaug.augmentables", "imgaug.augmentables.kps", "streamlit", "yacs",
                 "yacs.config", "ruamel", "ruamel.yaml", "seaborn", "inquirer", "streamlit_option_menu"]:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__getattr__ = _stub_getattr
            sys.modules[name] = m
    return True


def make_cfg(ncam=12, K=23, bbox=256, roi=144, spacing=2, size="small"):
    """The fields of the reference's yacs config that its inference classes read."""
    from types import SimpleNamespace as NS
    return NS(PARENT_DIR=REF, PROJECT_NAME="Example_Project", DATALOADER_NUM_WORKERS=0,
              DATASET=NS(DATASET_ROOT_DIR="datasets", DATASET_3D="Example_Dataset", DATASET_2D="Example_Dataset",
                         MEAN=[0.485, 0.456, 0.406], STD=[0.229, 0.224, 0.225], IMAGE_SIZE=[1280, 1024]),
              CENTERDETECT=NS(MODEL_SIZE=size, IMAGE_SIZE=256, NUM_JOINTS=1),
              KEYPOINTDETECT=NS(MODEL_SIZE=size, BOUNDING_BOX_SIZE=bbox, NUM_JOINTS=K),
              HYBRIDNET=NS(ROI_CUBE_SIZE=roi, GRID_SPACING=spacing, NUM_CAMERAS=ncam, BATCH_SIZE=1))
