import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # make sure the in-tree CUDA library exists (nvcc cross-compiles sm_100a without a GPU)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_msat_build", ROOT / "marl_sat_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except RuntimeError as e:          # no nvcc: use the prebuilt library if it is there
        if not mod.LIB_PATH.exists():
            raise pytest.UsageError(str(e))


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: GPU parity tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
