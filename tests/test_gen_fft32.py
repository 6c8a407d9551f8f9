"""The generated register-level 32-point DFT: op list checked with numpy, and the committed
header is what the generator emits."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
GEN = os.path.join(ROOT, "audioflux_b200", "csrc", "gen", "gen_fft32.py")


def _mod():
    spec = importlib.util.spec_from_file_location("gen_fft32", GEN)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_op_list_is_a_dft():
    m = _mod()
    rng = np.random.default_rng(5)
    for _ in range(4):
        x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
        assert np.abs(m.run_numpy(x) - np.fft.fft(x)).max() < 1e-12


def test_committed_header_is_current():
    m = _mod()
    path = os.path.join(ROOT, "audioflux_b200", "csrc", "kernels", "fft32_gen.cuh")
    assert open(path).read() == m.emit()
