"""GPU: per-kernel parity of libsfk (called through the C-ABI) against plain PyTorch fp32 on the same seeded inputs.
Each group runs tools/check_*.py in its own process so that a device trap in one kernel cannot poison the others.
Tolerances (relative to the reference tensor's max): fp32-out kernels 1e-4..1e-5, bf16-out kernels 1e-2..3e-2."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, group, timeout=600, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), group], cwd=ROOT, capture_output=True, text=True,
                       timeout=timeout, env=dict(os.environ, **(env or {})))
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0 and "ALLOK" in r.stdout, f"{script} {group} failed:\n{tail}"


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["nt", "nn", "tn", "epi", "v2"])
def test_gemm_tcgen05_layouts(group):
    """v2: every epilogue combination (bias, per-sample scale, fp32 residual through the coalesced transpose path, ReLU, bf16 /
    fp32 out, split-K reduce-add) on full, ragged and multi-column-tile shapes, B-stationary and streaming main loops."""
    _run("check_gemm.py", group)


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["nt", "nn", "tn", "epi", "v2"])
def test_gemm_cta_pair_layouts(group):
    """Same cases through the CTA-pair (tcgen05 cta_group::2, 2-CTA cluster) kernel, forced on for every eligible shape."""
    _run("check_gemm.py", group, env={"SFK_GEMM_CLUSTER2": "1"})


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["ln", "dwconv", "im2col", "attn_fwd", "attn_bwd", "decode", "loss", "optim"])
def test_ops(group):
    _run("check_ops.py", group)
