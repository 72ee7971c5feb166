"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), the host-side logic (seed padding,
block accounting, shard arithmetic), and the N>1 sharding path under gloo with
world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.binding.load()
    names = pkg.binding.declared_symbols()
    assert len(names) >= 80
    missing = [s for s in names if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.sclgpu_strerror(pkg.binding.EDETECT) == b"error detected during recovery"


def test_no_cpu_fallback_without_gpu(pkg):
    """On a box without a GPU the context refuses to come up (SCLGPU_ECUDA)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CudaError):
        pkg.Context(0)


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (test infrastructure)."""
    pdir = os.path.join(REPO, "secure-computation-library_b200")
    for root, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "scl_oracle" not in src, f
                assert "libscloracle" not in src and "libsclref" not in src, f
    for f in os.listdir(os.path.join(REPO, "include")):
        src = open(os.path.join(REPO, "include", f)).read()
        assert "scl_oracle" not in src and "sclo_" not in src and "sclref_" not in src, f


def test_seed_and_block_accounting(pkg):
    api, sh = pkg.api, pkg.sharding
    assert api.seed16("abc") == b"abc" + b"\0" * 13
    assert api.seed16(b"0123456789abcdefXYZ") == b"0123456789abcdef"  # prg.cc:88-101 truncation
    assert api.seed16(b"") == b"\0" * 16
    # B = ceil((t+1)*byteSize/16)  
    assert sh.blocks_per_share_call(61, 2) == 2
    assert sh.blocks_per_share_call(61, 15) == 8
    assert sh.blocks_per_share_call(127, 7) == 8
    assert sh.blocks_per_share_call(61, 0) == 1
    assert sh.blocks_for_random(61, 5) == 3 and sh.blocks_for_random(61, 5, True) == 5
    assert sh.blocks_for_random(127, 5) == 5
    v = api.from_ints([1, (1 << 127) - 2], 127)
    assert v.shape == (2, 2) and list(api.to_ints(v, 127)) == [1, (1 << 127) - 2]


def test_shard_range_properties(pkg):
    sh = pkg.sharding
    for N in (0, 1, 5, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 2):
                parts = [sh.shard_range(N, world, r, align) for r in range(world)]
                assert parts[0].lo == 0 and parts[-1].hi == N
                for a, b in zip(parts, parts[1:]):
                    assert a.hi == b.lo
                    assert a.hi % align == 0 or a.hi == N
                assert max(p.count for p in parts) - min(p.count for p in parts) <= align
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)
    s = sh.shard_range(1 << 26, 8, 3)
    assert sh.share_first_block(61, 15, 100, s) == 100 + 3 * (1 << 23) * 8
    assert sh.random_first_block(61, 7, sh.shard_range(100, 2, 1, 2)) == 7 + 25
    assert sh.random_first_block(127, 7, sh.shard_range(100, 2, 1)) == 7 + 50


def test_sharded_share_equals_unsharded_single_process(pkg, port):
    """Rank r sharing its slice with first_block + lo*B reproduces the slice of the
    one-PRG batch (the engine here is the oracle: this checks the HOST arithmetic)."""
    sh = pkg.sharding
    for field, t, n, N in [(61, 15, 32, 101), (127, 7, 16, 37), (61, 2, 5, 64)]:
        secrets = port.vector_random(field, "secrets", 0, N)
        full = port.shamir_share(field, secrets, t, n, "shamir bench", 11)
        for world in (2, 3, 8):
            parts = []
            for r in range(world):
                s = sh.shard_range(N, world, r)
                parts.append(port.shamir_share(field, secrets[s.lo:s.hi], t, n, "shamir bench",
                                               sh.share_first_block(field, t, 11, s)))
            assert np.array_equal(np.concatenate(parts, axis=0), full)
    for field in (61, 127):
        full = port.vector_random(field, "prg bench", 5, 1001)
        parts = []
        for r in range(4):
            s = sh.shard_range(1001, 4, r, align=2)
            parts.append(port.vector_random(field, "prg bench", sh.random_first_block(field, 5, s), s.count))
        assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_sharded_array_and_additive_share_equal_unsharded(pkg, port):
    """The same host arithmetic for array-valued sharings (ceil((t+1)*W*bs/16) blocks each) and additive
    sharings (n-1 blocks each)."""
    sh = pkg.sharding
    for field, W, t, n, N in [(61, 2, 15, 32, 45), (61, 3, 2, 5, 33), (127, 2, 7, 16, 21)]:
        es = () if field == 61 else (2,)
        secrets = port.vector_random(field, "pairs", 0, N * W).reshape((N, W) + es)
        full = port.shamir_share_array(field, secrets, t, n, "pedersen", 7)
        for world in (2, 5):
            parts = []
            for r in range(world):
                s = sh.shard_range(N, world, r)
                parts.append(port.shamir_share_array(field, secrets[s.lo:s.hi], t, n, "pedersen",
                                                     sh.array_share_first_block(field, W, t, 7, s)))
            assert np.array_equal(np.concatenate(parts, axis=0), full)
    for field, n, N in [(61, 4, 50), (127, 3, 17)]:
        secrets = port.vector_random(field, "secrets", 0, N)
        full = port.additive_share(field, secrets, n, "additive", 9)
        parts = []
        for r in range(3):
            s = sh.shard_range(N, 3, r)
            parts.append(port.additive_share(field, secrets[s.lo:s.hi], n, "additive", sh.additive_share_first_block(n, 9, s)))
        assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_gloo_world2_sharded_share_recover():
    """world_size-2 gloo run of the sharded driver (tests/dist_worker.py)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(REPO, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, env=env, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "DIST_OK world=2" in r.stdout, r.stdout[-3000:]


def test_multi_device_slicing_rule(pkg):
    """sclgpu_multi_slice (multi.cu; a pure function, no device): the slices of the multi-device handle tile [0, N)
    contiguously, start on multiples of `align`, differ in size by at most one aligned group, and agree with the
    per-process sharding rule (sharding.shard_range) -- so one process over G GPUs and G processes over one GPU each
    cut a batch, and its PRG stream, at the same places."""
    import ctypes as C

    lib, sh = pkg.binding.load(), pkg.sharding
    lo, hi = C.c_uint64(), C.c_uint64()
    for n_units in (0, 1, 2, 7, 64, 1000, 100003, (1 << 26) + 5):
        for parts in (1, 2, 3, 4, 8):
            for align in (1, 2):
                prev, sizes = 0, []
                for g in range(parts):
                    assert lib.sclgpu_multi_slice(n_units, parts, g, align, C.byref(lo), C.byref(hi)) == 0
                    assert lo.value == prev and hi.value >= lo.value and (lo.value % align == 0 or lo.value == n_units)
                    s = sh.shard_range(n_units, parts, g, align)
                    assert (s.lo, s.hi) == (lo.value, hi.value)
                    sizes.append(hi.value - lo.value)
                    prev = hi.value
                assert prev == n_units and max(sizes) - min(sizes) <= 2 * align
    assert lib.sclgpu_multi_slice(10, 0, 0, 1, C.byref(lo), C.byref(hi)) == pkg.binding.EINVAL
    assert lib.sclgpu_multi_slice(10, 2, 2, 1, C.byref(lo), C.byref(hi)) == pkg.binding.EINVAL


def test_async_and_multi_fail_loudly_without_gpu(pkg):
    """No device: the multi-device handle refuses to come up as a whole, and there is no context to be asynchronous on."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CudaError):
        pkg.MultiContext([0, 1])


def test_bitsliced_aes_host_build_vs_oracle():
    """csrc/aes_bitsliced.cuh compiled for the HOST (g++): the S-box circuit on all 256 inputs and 896 keystream blocks
    (four seeds, counters across the 2^32 boundary) against the oracle's util::PRG -- the same functions the GPU kernel
    k_prg_bitsliced runs (tests/cpp/bitsliced_check.cc)."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "tests", "cpp", "_build", "bitsliced_check")
    r = subprocess.run(["make", "-C", os.path.join(repo, "tests", "cpp"), "_build/bitsliced_check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "BITSLICED_CHECK PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
