"""The optional kernel variants stay parity-green: the per-kernel (layerwise), end-to-end training and inference parity
suites are re-run in a child process with the variant switched on.  The switches are read once per process / binding
(csrc/pcseg_api.cu), hence the child process.

  PCSEG_XF=1    transform-stage GEMMs: conv2 / conv3 / conv4 apply the train-mode BatchNorm + ReLU of their input on the A
                tiles in shared memory (pcs.py:106-109), no k_bn_relu launch; PCSEG_XF_SEG3=1 adds seg_conv3 with dropout
                (pcs.py:126-127)
  PCSEG_PAIR=1  cta_group::2 GEMMs (clusters of two CTAs, M = 256 per instruction) for global_feat forward / data gradient
                (pcs.py:113) and the inference max-pool GEMM (pcs.py:113-114)
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rerun(env_extra, files):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + [os.path.join(ROOT, "tests", f) for f in files]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert out.returncode == 0, (env_extra, out.stdout[-3000:], out.stderr[-2000:])
    assert " passed" in out.stdout and " failed" not in out.stdout, out.stdout[-1500:]


def test_transform_stage_gemms_pass_the_training_parity_suites():
    _rerun({"PCSEG_XF": "1", "PCSEG_XF_SEG3": "1"}, ["test_layerwise_gpu.py", "test_train_gpu.py"])


def test_cta_pair_gemms_pass_the_parity_suites():
    _rerun({"PCSEG_PAIR": "1"}, ["test_layerwise_gpu.py", "test_train_gpu.py", "test_eval_gpu.py"])


def test_cta_pair_and_transform_stage_together():
    _rerun({"PCSEG_PAIR": "1", "PCSEG_XF": "1"}, ["test_layerwise_gpu.py"])

