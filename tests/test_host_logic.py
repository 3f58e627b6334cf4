"""CPU-side tests of the product's host logic: C ABI surface, constructor validation, sharding.

No compute calls are made here (there is no GPU in the build container); validation runs
before any device is touched, so Config errors are observable on CPU.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import yagi_b200 as yb
from yagi_b200 import _lib
from yagi_b200.sharding import firpfbch2_time_shards, stream_shards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "yagi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(yg_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 45
    L = C.CDLL(_lib.path()) if os.path.exists(_lib.path()) else _lib.lib()
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    assert sorted(_lib.SYMBOLS) == declared          # the ctypes binding covers the whole header
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (yg_\w+)", out))
    assert exported == set(declared)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_error_string():
    L = _lib.lib()
    assert L.yg_version() == 0x000100
    n = C.c_int32(-1)
    assert L.yg_device_count(C.byref(n)) == 0 and n.value >= 0


def test_fir_design_kaiser_matches_formula_and_validates():
    h = yb.fir_design_kaiser(3585, 1.0 / 256, 60.0, 0.0)
    n = 3585
    beta = 0.1102 * (60.0 - 8.7)
    t = np.arange(n) - (n - 1) / 2.0
    r = 2 * t / (n - 1)
    ref = np.sinc(2 * t / 256) * np.i0(beta * np.sqrt(1 - r * r)) / np.i0(beta)
    assert np.abs(h - ref).max() < 5e-6
    for bad in [(0, 0.1, 60, 0), (10, 0.0, 60, 0), (10, 0.6, 60, 0), (10, 0.1, 0, 0), (10, 0.1, 60, -0.5), (10, 0.1, 60, 0.6)]:
        with pytest.raises(yb.ConfigError):
            yb.fir_design_kaiser(*bad)


def test_firpfbch2_crcf_config():
    """autotest firpfbch2_crcf_config through the product ABI (validation precedes device use)."""
    for bad in [(77, 76, 12), (yb.ANALYZER, 0, 12), (yb.ANALYZER, 17, 12), (yb.ANALYZER, 76, 0)]:
        with pytest.raises(yb.ConfigError):
            yb.FirPfbCh2.new_kaiser(bad[0], bad[1], bad[2], 60.0)
    with pytest.raises(yb.ConfigError):
        yb.FirPfbCh2.new(yb.ANALYZER, 8, 3, np.zeros(47, dtype=np.float32))
    with pytest.raises(yb.ConfigError):
        yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, 8, 3, 0.0)        # As must be > 0 (kaiser.rs:27-29)


def test_firpfbch_and_firfilt_config():
    for bad in [(77, 8, 4), (yb.ANALYZER, 0, 4), (yb.ANALYZER, 8, 0)]:
        with pytest.raises(yb.ConfigError):
            yb.FirPfbCh.new(bad[0], bad[1], bad[2], np.zeros(64, dtype=np.float32))
    with pytest.raises(yb.ConfigError):
        yb.FirPfbCh.new_kaiser(yb.ANALYZER, 8, 0, 60.0)
    with pytest.raises(yb.ConfigError):
        yb.FirPfbCh.new(yb.ANALYZER, 8, 4, np.zeros(31, dtype=np.float32))
    with pytest.raises(yb.ConfigError):
        yb.FirPfbCh.new(yb.ANALYZER, 8, 4, np.zeros(32, dtype=np.float32), n_streams=0)
    with pytest.raises(yb.ConfigError):
        yb.FirFilt.new(np.zeros(0, dtype=np.float32))                 # firfilt.rs:65-67
    with pytest.raises(yb.ConfigError):
        yb.FirFilt.new_kaiser(0, 0.25)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(yb.InternalError, match="no CPU fallback"):
        yb.FirPfbCh2.new_kaiser(yb.ANALYZER, 16, 5, 60.0)
    with pytest.raises(yb.InternalError):
        yb.FirFilt.new(np.ones(4, dtype=np.float32))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "yagi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


@pytest.mark.parametrize("n_frames,M,m,ws", [(2 ** 21, 256, 7, 8), (2 ** 21, 256, 7, 3), (1001, 16, 5, 4), (7, 8, 2, 8), (64, 1024, 4, 2)])
def test_time_shards_cover_and_align(n_frames, M, m, ws):
    sh = firpfbch2_time_shards(n_frames, M, m, ws)
    assert len(sh) == ws
    assert sh[0].frame_begin == 0 and sh[-1].frame_end == n_frames
    for a, b in zip(sh, sh[1:]):
        assert a.frame_end == b.frame_begin
    for s in sh:
        assert s.frame_begin % 2 == 0                      # local parity == global parity
        assert s.sample_begin == s.frame_begin * M // 2
        assert s.halo_len == (4 * m - 1) * M // 2 == 2 * M * m - M // 2
        assert s.halo_begin == s.sample_begin - s.halo_len
    sizes = [s.n_frames for s in sh]
    assert max(sizes) - min(sizes) <= 3


def test_stream_shards():
    sh = stream_shards(4096, 8)
    assert [len(r) for r in sh] == [512] * 8
    sh = stream_shards(10, 4)
    assert sum(len(r) for r in sh) == 10 and sh[0].start == 0 and sh[-1].stop == 10


def test_numa_helper_parses_cpulists_and_is_a_no_op_without_a_gpu():
    """yagi_b200/_numa.py: cpulist parsing; without a CUDA device (or without sysfs NUMA data) nothing is changed."""
    import os
    from yagi_b200._numa import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    import torch
    if not torch.cuda.is_available():
        assert bind_to_gpu_numa_node(0) is None
        assert os.sched_getaffinity(0) == before


def test_rust_sys_crate_tracks_the_c_abi():
    """rust/yagi-b200-sys cannot be compiled here (no cargo), so keep it honest textually:
    build.rs compiles every csrc/*.cu (globbed, no hand-written list to go stale) and src/lib.rs
    declares exactly the symbols of include/yagi_b200.h with the same number of arguments."""
    hdr = open(os.path.join(ROOT, "include", "yagi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = {}
    for name, args in re.findall(r"\b(yg_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = args.strip()
        declared[name] = 0 if args in ("", "void") else args.count(",") + 1
    rs = open(os.path.join(ROOT, "rust", "yagi-b200-sys", "src", "lib.rs")).read()
    bound = {}
    for name, args in re.findall(r"pub fn (yg_[A-Za-z0-9_]+)\s*\(([^)]*)\)", rs):
        args = args.strip()
        bound[name] = 0 if args == "" else args.count(":")
    assert bound == declared
    build_rs = open(os.path.join(ROOT, "rust", "yagi-b200-sys", "build.rs")).read()
    assert "read_dir(&csrc)" in build_rs and 'x == "cu"' in build_rs          # globbed
    for f in os.listdir(os.path.join(ROOT, "yagi_b200", "csrc")):
        if f.endswith(".cu"):
            assert ('"%s"' % f[:-3]) not in build_rs                             # no stale explicit list
    # the facades call only declared symbols
    for facade in ("multichannel/mod.rs", "filter/firfilt_gpu.rs"):
        text = open(os.path.join(ROOT, "rust", facade)).read()
        used = set(re.findall(r"sys::(yg_[A-Za-z0-9_]+)\s*\(", text))
        assert used and used <= set(declared), used - set(declared)


def test_mixed_radix_stockham_index_formulas():
    """The index arithmetic of the generic kernels' transform (common.cuh: block_dft / slot_dft / tiled_pass), restated
    in numpy: one pass per prime factor (4 before 2, then odd primes ascending); a pass of radix r maps butterfly
    j = jh Ns + k from x[j + i M/r] to y[(jh r + q) Ns + k] with the single table exponent i (k + q Ns) M / (Ns r),
    which stays below M for the pass twiddle alone (i k M / (Ns r)).  Checked against numpy's FFT for composite, prime
    and power-of-two lengths."""
    import numpy as np

    def radices(M):
        out, rem = [], M
        while rem > 1:
            r = rem
            if rem % 4 == 0:
                r = 4
            elif rem % 2 == 0:
                r = 2
            else:
                p = 3
                while p * p <= rem:
                    if rem % p == 0:
                        r = p
                        break
                    p += 2
            out.append(r)
            rem //= r
        return out

    def transform(xin):
        M = len(xin)
        tw = np.exp(2j * np.pi * np.arange(M) / M)
        x, Ns = xin.astype(np.complex128), 1
        for r in radices(M):
            L, step = M // r, M // (Ns * r)
            y = np.empty(M, dtype=np.complex128)
            for o in range(M):
                k, t = o % Ns, o // Ns
                q, jh = t % r, t // r
                j = jh * Ns + k
                e = (k + q * Ns) * step
                assert e < M and (r - 1) * k * step < M
                idx = (np.arange(r) * e) % M
                y[o] = np.sum(x[j + np.arange(r) * L] * tw[idx])
            x, Ns = y, Ns * r
        return x

    rng = np.random.default_rng(3)
    assert radices(1000) == [4, 2, 5, 5, 5] and radices(48) == [4, 4, 3] and radices(202) == [2, 101]
    for M in (2, 6, 10, 12, 24, 48, 64, 66, 100, 126, 202, 240, 250, 384, 1000):
        x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
        ref = np.fft.ifft(x) * M
        assert np.abs(transform(x) - ref).max() <= 1e-11 * np.abs(ref).max(), M
