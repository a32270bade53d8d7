"""CPU: the C-ABI library loads, exports every symbol declared in include/sfk.h, and refuses to run without an sm_100 GPU."""
import ctypes
import os

import pytest
import torch

from semisupervisedobjectdetection_b200 import _lib as L
from semisupervisedobjectdetection_b200 import ops


def test_library_exports_every_declared_symbol():
    lib = L.load()
    protos = L.header_prototypes()
    assert len(protos) >= 50
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/sfk.h but not exported by libsfk.so"
    assert lib.sfk_version() == 1
    assert isinstance(lib.sfk_last_error(), bytes)


def test_struct_mirrors_match_header_field_order():
    with open(L.HEADER_PATH) as f:
        text = f.read()
    import re
    body = re.search(r"typedef struct \{([^}]*)\} sfk_layer_params;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*([a-z0-9_]+)", body)
    assert tuple(names) == ops._PARAM_FIELDS
    body = re.search(r"typedef struct \{([^}]*)\} sfk_layer_grads;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*([a-z0-9_]+)", body)
    assert tuple(names) == ops._GRAD_FIELDS
    assert ctypes.sizeof(ops.LayerDims) == 8 * 4
    assert ctypes.sizeof(ops.LayerParams) == len(ops._PARAM_FIELDS) * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = L.load()
    assert lib.sfk_device_check() != 0
    with pytest.raises(L.SfkError):
        L.ensure_device()
    from semisupervisedobjectdetection_b200.configuration import SegformerConfig
    from semisupervisedobjectdetection_b200.modeling import SegformerForSemanticSegmentation
    model = SegformerForSemanticSegmentation(SegformerConfig(num_labels=1)).eval()
    with pytest.raises(L.SfkError):
        model(pixel_values=torch.rand(1, 3, 64, 64))
    from semisupervisedobjectdetection_b200.loss import SegmentationLoss
    with pytest.raises((L.SfkError, RuntimeError)):
        SegmentationLoss(1, loss_type="dice", activation="none")(torch.rand(2, 8, 8), torch.ones(2, 8, 8))


def test_workspace_size_queries_run_on_cpu():
    lib = L.load()
    d = ops.LayerDims(16, 32, 32, 320, 5, 2, 1280, 0)
    saved = lib.sfk_layer_saved_bytes(ctypes.byref(d))
    scratch = lib.sfk_layer_scratch_bytes(ctypes.byref(d))
    tokens = 16 * 32 * 32
    assert saved > tokens * 320 * 2 * 3 and saved < tokens * 320 * 64
    assert scratch > saved // 4
    assert lib.sfk_dice_workspace_bytes(16) > 0 and lib.sfk_loss_workspace_bytes() > 0
