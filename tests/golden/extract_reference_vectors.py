#!/usr/bin/env python
"""Extract the reference's golden vectors for the channelizer's building blocks.

Reads the `pub const NAME: [T; N] = [...]` arrays out of the reference's in-file
test data (read-only, /root/reference) and writes them into one compressed .npz
fixture next to this script, so that tests can pin the oracle (and, through the
oracle-free delta-prototype trick, the CUDA FFT stage) on machines where
/root/reference does not exist (the GPU box).

Sources (all under /root/reference/src):
  fft/test_data.rs                      33 X/Y pairs for Fft::run       (fft/mod.rs:156-352)
  filter/fir/firfilt_test_data.rs       crcf h/x/y triplets             (firfilt.rs:961-999)
  filter/fir/firdecim_test_data.rs      crcf h/x/y triplets             (firdecim.rs:355-414)
  dotprod/mod.rs:455-524                crcf rand01 / rand02 known answers
  filter/fir/firpfb.rs:310-359          48-tap 4-branch impulse-response known answers

Run:  python tests/golden/extract_reference_vectors.py
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("YAGI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "reference_vectors.npz")

_CONST = re.compile(r"const (\w+): \[([\w<>]+); (\d+)\] = \[(.*?)\];", re.S)
_CPLX = re.compile(r"new\(\s*([-+\d.eE]+),\s*([-+\d.eE]+)\s*\)")
_REAL = re.compile(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?")


def parse_consts(path):
    out = {}
    text = open(path).read()
    for name, ty, n, body in _CONST.findall(text):
        n = int(n)
        if "Complex" in ty:
            vals = np.array([complex(float(a), float(b)) for a, b in _CPLX.findall(body)], dtype=np.complex128)
        else:
            body = re.sub(r"//.*", "", body)
            vals = np.array([float(v) for v in _REAL.findall(body)], dtype=np.float64)
        assert vals.size == n, (path, name, vals.size, n)
        out[name] = vals
    return out


def parse_let_array(text, varname, start=0):
    """`let NAME: [T; N] = [ ... ];` inside a test function body."""
    m = re.compile(r"let %s: \[(\w+); (\d+)\] = \[(.*?)\];" % varname, re.S).search(text, start)
    ty, n, body = m.group(1), int(m.group(2)), m.group(3)
    body = re.sub(r"//.*", "", body)
    if ty.startswith("Cf32") or "Complex" in ty:
        vals = np.array([complex(float(a), float(b)) for a, b in _CPLX.findall(body)], dtype=np.complex128)
    else:
        vals = np.array([float(v) for v in _REAL.findall(body)], dtype=np.float64)
    assert vals.size == n, (varname, vals.size, n)
    return vals, m.end()


def collect(ref=REF):
    src = os.path.join(ref, "src")
    out = {}
    for k, v in parse_consts(os.path.join(src, "fft", "test_data.rs")).items():
        out[k] = v
    for k, v in parse_consts(os.path.join(src, "filter", "fir", "firfilt_test_data.rs")).items():
        if "_CRCF_" in k:
            out[k] = v
    for k, v in parse_consts(os.path.join(src, "filter", "fir", "firdecim_test_data.rs")).items():
        if "_CRCF_" in k:
            out[k] = v

    # dotprod crcf rand01 / rand02 (inline data)
    text = open(os.path.join(src, "dotprod", "mod.rs")).read()
    for tag in ("rand01", "rand02"):
        at = text.index("fn test_dotprod_crcf_%s" % tag)
        h, e = parse_let_array(text, "h", at)
        x, e = parse_let_array(text, "x", e)
        t = re.compile(r"let test = Cf32::new\(\s*([-+\d.eE]+),\s*([-+\d.eE]+)\)").search(text, e)
        out["DOTPROD_CRCF_%s_H" % tag.upper()] = h
        out["DOTPROD_CRCF_%s_X" % tag.upper()] = x
        out["DOTPROD_CRCF_%s_Y" % tag.upper()] = np.array([complex(float(t.group(1)), float(t.group(2)))])
        if tag == "rand01":
            t = re.compile(r"let test_rev = Cf32::new\(\s*([-+\d.eE]+),\s*([-+\d.eE]+)\)").search(text, e)
            out["DOTPROD_CRCF_RAND01_YREV"] = np.array([complex(float(t.group(1)), float(t.group(2)))])

    # firpfb impulse-response known answers
    text = open(os.path.join(src, "filter", "fir", "firpfb.rs")).read()
    at = text.index("fn test_firpfb_impulse_response")
    h, e = parse_let_array(text, "h", at)
    noise, e = parse_let_array(text, "noise", e)
    test, e = parse_let_array(text, "test", e)
    out["FIRPFB_IMPULSE_H"] = h
    out["FIRPFB_IMPULSE_NOISE"] = noise
    out["FIRPFB_IMPULSE_TEST"] = test
    return out


def main():
    if not os.path.isdir(REF):
        print("reference not found at", REF, file=sys.stderr)
        return 1
    vecs = collect()
    np.savez_compressed(OUT, **vecs)
    print("wrote %s: %d arrays, %d bytes" % (OUT, len(vecs), os.path.getsize(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
