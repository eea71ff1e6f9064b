"""CPU tests: host-side graph model, C-ABI surface, oracle self-consistency."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import _lib, alist, harness
from oracle import binding as ob

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    declared = set(re.findall(r"\b(ldpc_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ldpc_b200_version()


def test_ctypes_struct_layouts_match_the_library():
    """The Python mirror of the four public structs has the sizes the library was compiled with."""
    import ctypes as C

    sizes = (C.c_uint32 * 4)()
    _lib.lib().ldpc_b200_abi_sizes(sizes)
    assert list(sizes) == [C.sizeof(_lib.CodeInfo), C.sizeof(_lib.Config), C.sizeof(_lib.DecodeParams),
                           C.sizeof(_lib.Report)]


@pytest.mark.parametrize("name", ["reg36", "readme13", "hideg", "tiny"])
def test_graph_tables_match_oracle_and_generator(small_codes, name):
    gen, path = small_codes[name]
    c = L.Code(path=path)
    o = ob.OracleCode(path=path)
    assert (c.n_checks, c.n_vars, c.n_edges, c.n_erased) == (o.n_checks, o.n_vars, o.n_edges, o.n_erased)
    assert (c.max_deg_var, c.max_deg_check) == (o.max_deg_var, o.max_deg_check)
    for attr in ("row_ptr", "col_ptr", "in_to_out", "out_to_in", "out_edge_to_var"):
        assert np.array_equal(getattr(c, attr), getattr(o, attr)), attr
    col_ptr, in_to_out = gen.csc()
    assert np.array_equal(c.col_ptr, col_ptr) and np.array_equal(c.in_to_out, in_to_out)
    assert np.array_equal(c.out_edge_to_var, gen.col_idx)
    assert c.rate() == pytest.approx(o.rate())
    # CSR constructor gives the same tables as the alist route
    c2 = L.Code.from_generated(gen)
    assert np.array_equal(c2.in_to_out, c.in_to_out) and np.array_equal(c2.col_ptr, c.col_ptr)


def test_alist_dialect_details():
    # checks first, line 2 ignored, "#e=" / "#ec=" headers, zero padding on rows (ldpc_code.cpp:52-151)
    text = "#e=32\n#ec=1\n#foo=bar\n2 64\n99 99\n" + "33 33\n" + " ".join(["1"] * 62 + ["2", "2"]) + "\n"
    rows = [list(range(1, 33)) + [63], list(range(33, 65)) + [64]]
    rows[0][-1], rows[1][-1] = 63, 64
    rows = [list(range(1, 32)) + [63, 64], list(range(32, 63)) + [63, 64]]
    text += " ".join(map(str, rows[0])) + " 0 0\n" + " ".join(map(str, rows[1])) + "\n"
    c = L.Code(text=text)
    o = ob.OracleCode(text=text)
    assert (c.n_checks, c.n_vars, c.n_edges, c.n_erased, c.n_erased_checks) == (2, 64, 66, 32, 1)
    assert np.array_equal(c.in_to_out, o.in_to_out)
    # variable 63 (index 62) sits in both checks: in-edges ordered by increasing check index
    a = c.col_ptr[62]
    assert list(c.in_to_out[a : a + 2]) == [31, 64]


def test_error_paths():
    with pytest.raises(L.LdpcError) as e:
        L.Code(path="/nonexistent/file.alist")
    assert e.value.status == _lib.ERR_IO and "could not be opened" in str(e.value)
    with pytest.raises(L.LdpcError) as e:
        L.Code(text="2 4\n2 1\n2 2\n1 1 1 2\n1 2\n3 4\n")  # degree sums differ
    assert e.value.status == _lib.ERR_ALIST
    with pytest.raises(ValueError):
        ob.OracleCode(text="2 4\n2 1\n2 2\n1 1 1 2\n1 2\n3 4\n")
    # N not a multiple of 32 is rejected before any device work (ldpc_decoder_gpu.cu:30-32)
    g = alist.make_regular_code(96, 3, 6, seed=1)
    bad = alist.Code(g.n_checks, 100, g.row_ptr, g.col_idx)
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(L.Code.from_generated(bad))
    assert e.value.status == _lib.ERR_MULTIPLE_32
    # a variable without edges: "Incorrect code structure" (ldpc_decoder_gpu.cu:40-58)
    bad2 = alist.Code(g.n_checks, 128, g.row_ptr, g.col_idx)
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(L.Code.from_generated(bad2))
    assert e.value.status == _lib.ERR_CODE


def test_no_cpu_fallback_without_device(small_codes):
    from conftest import has_gpu

    if has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(L.Code(path=small_codes["tiny"][1]))
    assert e.value.status == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_syndrome_helpers_agree(small_codes):
    gen, path = small_codes["readme13"]
    o = ob.OracleCode(path=path)
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2**32, size=(5, gen.n_vars // 32), dtype=np.uint32)
    assert np.array_equal(gen.syndrome_packed(bits), o.syndrome(bits))


def test_oracle_phi_properties():
    x = np.array([0.0, 1e-6, 1e-5, 1e-3, 0.1, 0.5, 1.0, 2.0, 4.999, 5.0, 5.001, 10.0, 50.0], dtype=np.float32)
    y = ob.phi_abs(x)
    assert y[0] == y[1] == y[2] == pytest.approx(np.log(2e5), rel=1e-5)  # clamp at 1e-5 (flood.cu:14)
    ref = -np.log(np.tanh(np.maximum(x.astype(np.float64), 1e-5) / 2))
    small = x <= 5
    assert np.allclose(y[small], ref[small], rtol=1e-5)  # float log of a ratio near 1 costs ~5e-6 at x~5
    assert np.allclose(y[~small], 2 * np.exp(-x[~small].astype(np.float64)), rtol=2e-6)  # Taylor branch
    # phi is an involution on the unclamped range
    z = ob.phi_abs(ob.phi_abs(np.array([0.01, 0.3, 1.0, 3.0], dtype=np.float32)))
    assert np.allclose(z, [0.01, 0.3, 1.0, 3.0], rtol=1e-4)
    # fp16 table: monotone non-increasing, clamp plateau below 0x3f
    t = ob.phi_table_h().view(np.float16).astype(np.float32)
    assert (t[:0x40] == t[0x3F]).all()
    # non-increasing on each branch; the switch to 2e^-x at x = 5 (0x4500) jumps up slightly, as in the reference
    assert (np.diff(t[0x3F:0x4501]) <= 0).all() and (np.diff(t[0x4501:0x7C00]) <= 0).all()


@pytest.mark.parametrize("precision", [ob.FP32, ob.FP16])
def test_oracle_decodes_and_schedules_like_the_reference(small_codes, precision):
    gen, path = small_codes["reg36"]
    o = ob.OracleCode(path=path)
    n_frames, log2p = 20, 3
    dt = np.float32 if precision == ob.FP32 else np.float16
    inp, syn, bits = harness.make_frames(gen, n_frames, "awgn", 0.70, seed=4, dtype=dt)
    dec = ob.OracleDecoder(o, precision, log2p)
    res, iters, rep = dec.decode(ob.AWGN, harness.channel_param("awgn", 0.70), 50, 5, inp, syn)
    assert harness.count_bit_errors(res, bits).sum() == 0
    # first batch counts g+1 (iter_start = -1u), refilled frames multiples of k (ldpc_decoder_gpu.cu:306-309, 396)
    assert set(iters[:8] % 5) == {1} and set(iters[8:] % 5) == {0}
    assert rep["max_iter"] == iters.max() and rep["min_iter"] == iters.min()
    # decoded words satisfy the target syndromes
    assert np.array_equal(gen.syndrome_packed(res), syn)


def test_oracle_kernel_steps_equal_its_decode(small_codes):
    gen, path = small_codes["readme13"]
    o = ob.OracleCode(path=path)
    inp, syn, bits = harness.make_frames(gen, 4, "awgn", 0.5, seed=2)
    n_reg = gen.n_vars - gen.n_erased
    dec = ob.OracleDecoder(o, ob.FP32, 2)
    dec.refill(ob.AWGN, harness.channel_param("awgn", 0.5), inp[:n_reg], syn)
    pri = dec.priors(1)
    assert np.allclose(pri[:n_reg], inp[:n_reg, 1] * np.float32(harness.channel_param("awgn", 0.5)))
    assert (pri[n_reg:] == 0).all()
    for it in range(30):
        dec.backward()
        dec.forward(final_bits=True)
        if not dec.check_parity().any():
            break
    assert it < 29
    assert np.array_equal(dec.deinterlace(), bits)


def test_cpp_dropin_class_compiles_and_reports_errors(tmp_path):
    """The C++14 caller in tests/cpp_dropin_test.cpp builds against the drop-in header with -std=c++14 and the
    reference's exception type comes back for an unreadable alist (no GPU needed for that path)."""
    import subprocess

    exe = str(tmp_path / "cpp_dropin_test")
    libdir = os.path.join(ROOT, "ldpc_decoder_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp_dropin_test.cpp"), "-o", exe, "-L" + libdir, "-lldpc_b200",
                           "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, "/nonexistent.alist"], capture_output=True, text=True)
    assert r.returncode == 5 and "could not be opened" in r.stderr


def test_fast_cpu_baseline_kernels_agree_with_exact_oracle(small_codes):
    """The SIMD-over-frames kernels timed as the CPU baseline compute the same thing as the exact oracle."""
    gen, path = small_codes["readme13"]
    o = ob.OracleCode(path=path)
    inp, syn, bits = harness.make_frames(gen, 40, "awgn", 0.6, seed=3)
    prm = harness.channel_param("awgn", 0.6)
    n_reg = gen.n_vars - gen.n_erased
    exact = ob.OracleDecoder(o, ob.FP32, 4, pre_threshold=ob.PRE_THRESHOLD_OPENCL)
    fast = ob.OracleDecoder(o, ob.FP32, 4, pre_threshold=ob.PRE_THRESHOLD_OPENCL, fast=True)
    for d in (exact, fast):
        d.refill(ob.AWGN, prm, inp[:n_reg, :16], syn[:16])
        for _ in range(3):
            d.backward()
            d.forward(True)
    a, b = exact.messages(5), fast.messages(5)
    assert np.allclose(a, b, rtol=2e-4, atol=1e-6)
    assert np.array_equal(exact.deinterlace(), fast.deinterlace())
    r1, i1, _ = exact.decode(ob.AWGN, prm, 60, 10, inp, syn, simd_pairs=True)
    r2, i2, _ = fast.decode(ob.AWGN, prm, 60, 10, inp, syn, simd_pairs=True)
    assert np.array_equal(i1, i2) and np.array_equal(r1, r2)


def test_frame_queue_object_and_shared_memory_cursor():
    """The shared frame queue (include/ldpc_b200.h) needs no device: private and POSIX-shared-memory flavours, sizes,
    a second handle on the same shared object, argument checks."""
    import os

    import ldpc_decoder_b200 as L

    q = L.FrameQueue(1000, 32)
    assert q.size == 1000
    q.reset()
    q.close()
    name = f"ldpc_b200_test_{os.getpid()}"
    owner = L.FrameQueue(4096, 64, shm_name=name, create=True)
    other = L.FrameQueue(shm_name=name, create=False)
    assert other.size == 4096
    other.close()
    owner.close()
    with pytest.raises(L.LdpcError):
        L.FrameQueue(shm_name=name, create=False)       # unlinked by its owner
    with pytest.raises(L.LdpcError):
        L.FrameQueue(0, 0)


def test_locality_storage_order_recovers_hidden_locality():
    """The in-decoder storage order (LDPC_B200_ORDER_LOCALITY) is a permutation of checks and of
    variables, and on a banded code whose rows and columns were shuffled it brings the message rows a variable gathers
    back together (mean span of a variable's rows, in storage rows)."""
    import ldpc_decoder_b200 as L
    from ldpc_decoder_b200 import alist

    rng = np.random.default_rng(5)
    n, win = 8192, 96
    sockets = np.repeat(np.arange(n, dtype=np.int64), 3)
    for b in range(0, sockets.size, win):
        blk = sockets[b:b + win]
        sockets[b:b + win] = blk[rng.permutation(blk.size)]
    m = n // 2
    band = alist.Code(m, n, np.arange(0, 6 * m + 1, 6, dtype=np.uint32), sockets.astype(np.uint32), 0)
    vperm, cperm = rng.permutation(n), rng.permutation(m)
    rows = [vperm[band.col_idx[band.row_ptr[c]:band.row_ptr[c + 1]]] for c in cperm]
    gen = alist.Code(m, n, band.row_ptr, np.concatenate(rows).astype(np.uint32), 0)
    co, vo = L.Code.from_generated(gen).locality_order()
    assert np.array_equal(np.sort(co), np.arange(m)) and np.array_equal(np.sort(vo), np.arange(n))

    def mean_span(check_order):
        pos = np.empty(m, dtype=np.int64)
        pos[check_order] = np.arange(m)
        check_of_edge = np.repeat(pos, 6)                 # storage position of each edge's check (6 rows per check)
        order = np.argsort(gen.col_idx, kind="stable")
        per_var = check_of_edge[order].reshape(n, 3)
        return float((per_var.max(axis=1) - per_var.min(axis=1)).mean()) * 6

    before, after = mean_span(np.arange(m)), mean_span(co.astype(np.int64))
    assert after < before / 20, (before, after)


def test_dropin_half_conversions_match_numpy(tmp_path):
    """include/ldpc_b200_dropin.hpp converts halves on the host for channels without a device kernel (C++14 has no half
    type): float -> half rounds to nearest even exactly like numpy, incl. subnormals, ties, overflow, infinities and NaN;
    half -> float is exact for all 65536 patterns."""
    import subprocess

    exe = str(tmp_path / "cpp_half_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp_half_test.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "ldpc_decoder_b200"), "-lldpc_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "ldpc_decoder_b200")])
    rng = np.random.default_rng(3)
    halves = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    ties = (halves[:-1].astype(np.float64) + halves[1:].astype(np.float64)) / 2           # exact midpoints of neighbours
    vals = np.concatenate([
        np.exp(rng.uniform(np.log(1e-9), np.log(1e6), 200000)).astype(np.float32) * rng.choice([-1, 1], 200000).astype(np.float32),
        ties[np.isfinite(ties)].astype(np.float32),
        halves[np.isfinite(halves)],
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65519.99, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8],
                 dtype=np.float32)])
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    vals.tofile(src)
    subprocess.check_call([exe, src, dst])
    raw = np.fromfile(dst, dtype=np.uint8)
    got = raw[: 2 * vals.size].view(np.uint16)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    nan = np.isnan(vals)
    assert np.array_equal(got[~nan], want[~nan])
    assert np.all((got[nan] & 0x7C00) == 0x7C00) and np.all(got[nan] & 0x3FF)
    back = raw[2 * vals.size:].view(np.float32)
    ok = ~np.isnan(halves)
    assert np.array_equal(back[ok].view(np.uint32), halves[ok].view(np.uint32)) and np.all(np.isnan(back[~ok]))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (absent on the GPU box)")
def test_reference_main_cpp_compiles_and_links_against_the_dropin_header():
    """Build-only half of the drop-in claim (the GPU half is test_reference_main_cpp_builds_and_runs_against_the_dropin_header):
    the reference's own src/main.cpp and harness sources, unchanged, compile with include/ of this repository first on
    the include path and link with libldpc_b200.so (oracle/Makefile refdropin); the program starts and prints the
    reference's usage text without a device."""
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refdropin"])
    for prec in ("fp16", "fp32"):
        exe = os.path.join(ROOT, "oracle", "_ref", f"dropin_ldpc_decoder_{prec}")
        r = subprocess.run([exe, "-h"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and "-f s where s is the name of the code file" in r.stdout
        ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libldpc_b200.so" in ldd and "libcudart" not in ldd.split("libldpc_b200.so")[0]
