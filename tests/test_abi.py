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


def test_ctypes_structs_match_the_header(tmp_path):
    """The descriptor structs of include/pqlb200.h and their ctypes mirrors in pql_b200/_lib.py must agree
    in size and field offsets (a silent mismatch would hand the kernels garbage pointers)."""
    import ctypes as C
    import shutil
    import subprocess
    cxx = shutil.which("g++") or shutil.which("gcc")
    if cxx is None:
        pytest.skip("no host compiler")
    probes = {"pqlb_gemm_desc": (_lib.GemmDesc, ["M", "tile_n", "splits", "cluster", "noise_bound", "g"]),
              "pqlb_gemm_group": (_lib.GemmGroup, ["a", "a2", "bias", "q", "out2", "split_stride"]),
              "pqlb_mlp_desc": (_lib.MlpDesc, ["M", "n_groups", "g"]),
              "pqlb_mlp_group": (_lib.MlpGroup, ["x", "w2", "q", "h3", "act_w", "act_ldo", "noise_std", "act_n"]),
              "pqlb_mlp_bwd_desc": (_lib.MlpBwdDesc, ["M", "g"]),
              "pqlb_wgrad_desc": (_lib.WgradDesc, ["K", "n_problems", "p"]),
              "pqlb_wgrad_problem": (_lib.WgradProblem, ["dz", "h", "part", "split_stride", "M", "tile_n", "splits"]),
              "pqlb_colsum_desc": (_lib.ColsumDesc, ["n", "rows", "dz", "ld", "n_cols", "part"]),
              "pqlb_dp_desc": (_lib.DpDesc, ["grad_peers", "red_peers", "ctl_peers", "local", "rank", "world", "grid"])}
    lines = ['#include "pqlb200.h"', "#include <cstdio>", "#include <cstddef>", "int main() {"]
    for cname, (_, fields) in probes.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines.append("  return 0; }")
    src = tmp_path / "layout.cpp"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cxx, "-x", "c++", "-I", os.path.join(entry.ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        ctype, fields = probes[name]
        assert C.sizeof(ctype) == int(size), name
        assert [getattr(ctype, f).offset for f in fields] == [int(o) for o in offs], name
