"""The committed fixtures under tests/golden/ are recordings of the UNMODIFIED reference (oracle/make_golden*.py).
Where the reference tree exists (the build container) this test records them again into a scratch directory and
requires bit-equality; on the GPU box (no /root/reference) it is skipped."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="needs /root/reference")


def _same(a_path, b_path):
    a, b = np.load(a_path, allow_pickle=True), np.load(b_path, allow_pickle=True)
    assert sorted(a.files) == sorted(b.files), (a_path, sorted(a.files), sorted(b.files))
    for k in a.files:
        x, y = a[k], b[k]
        assert x.dtype == y.dtype and x.shape == y.shape, (a_path, k)
        if x.dtype == object:
            assert x.tolist() == y.tolist(), (a_path, k)
        else:
            assert x.tobytes() == y.tobytes(), (a_path, k)     # bit for bit (NaN-safe)


@pytest.mark.parametrize("module,prefix", [("oracle.make_golden", "dps_"), ("oracle.make_golden_psld", "psld_"),
                                           ("oracle.make_golden_resample", "resample_"),
                                           ("oracle.make_golden_pgdm", "pgdm_")])
def test_sampler_recordings_are_reproduced_bit_for_bit(tmp_path, monkeypatch, module, prefix):
    mod = importlib.import_module(module)
    monkeypatch.setattr(mod, "OUT", str(tmp_path))
    threads = torch.get_num_threads()
    torch.set_num_threads(1)          # as the generators do: a fixed reduction order
    try:
        for name, cfg in mod.CASES.items():
            mod.run_case(name, cfg)
    finally:
        torch.set_num_threads(threads)
    made = sorted(os.listdir(tmp_path))
    assert made and all(f.startswith(prefix) for f in made), made
    for f in made:
        _same(os.path.join(tmp_path, f), os.path.join(GOLDEN, f))
    committed = sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix))
    assert [f for f in committed if f not in made] == [], "a committed fixture no generator produces"


def test_io_recordings_are_reproduced_bit_for_bit(tmp_path, monkeypatch):
    mod = importlib.import_module("oracle.make_golden_io")
    monkeypatch.setattr(mod, "OUT", str(tmp_path))
    mod.main()
    for f in sorted(os.listdir(tmp_path)):
        _same(os.path.join(tmp_path, f), os.path.join(GOLDEN, f))
