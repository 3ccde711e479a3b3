"""Runs the C++ host-layer parity tests (tests/cpp/test_tensor_op_b200.cpp: TensorOpB200 behind the reference's
TensorOp interface vs the oracle, written like Neuro.Tests/src/TensorOpGpuTests.cpp) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_tensor_op_b200")


def _build():
    import __graft_entry__ as graft
    graft.build()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_cpp_host_layer(mode):
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE, mode], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout


def test_cpp_host_layer_builds():
    """CPU check: the header-only host layer compiles and links against the C ABI."""
    _build()
    assert os.path.exists(EXE)
    out = subprocess.check_output(["nm", "-D", "--undefined-only", EXE]).decode()
    for sym in ("nb200_conv2d_forward", "nb200_conv2d_input_gradient", "nb200_conv2d_kernels_gradient", "nb200_adam_step"):
        assert sym in out
