"""The native extensions exist, import on a CPU-only box, and the CUDA objects really are sm_100a code
using the Blackwell paths (SASS mnemonics, B200_PROFILING.md §"What proves a Blackwell-native kernel")."""
import shutil
import subprocess

import pytest

from pytorch_ps_mpi_b200.ops import ext


def test_host_extension_imports_and_works():
    h = ext.host()
    assert h.ANY_SOURCE == -1
    raw = bytes(range(64))
    assert bytes(h.byteunshuffle(h.byteshuffle(raw, 4), 4)) == raw


def test_cuda_extension_imports_without_a_gpu():
    if not ext.CUDA_SO.exists():
        pytest.skip("CUDA extension not built yet (run __graft_entry__.build())")
    m = ext.cuda()
    for name in ("SymmBlock", "UpdatePlan", "encode", "signal", "wait_flags", "select_ready", "bcast_gemm",
                 "bn_forward", "bn_backward", "maxpool_forward", "im2col_stem", "normalize_nhwc3", "launch_count"):
        assert hasattr(m, name), name
    assert m.TILE == 2048 and m.MAX_RANKS >= 8


def _sass(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    return subprocess.run([exe, "-sass", str(obj)], stdout=subprocess.PIPE, text=True, check=True).stdout


def test_sass_shows_blackwell_paths():
    gemm, gemm2, psk = ext.OBJ / "bcast_gemm.o", ext.OBJ / "bcast_gemm2.o", ext.OBJ / "ps_kernels.o"
    if not gemm.exists() or not psk.exists() or not gemm2.exists():
        pytest.skip("object files not present (built elsewhere)")
    s = _sass(gemm) + _sass(gemm2)
    assert "sm_100a" in s or "SM100" in s.upper() or "EF_CUDA_SM100" in s
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.2D.2CTA", "LDTM", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in s, f"{mnemonic} missing from bcast_gemm SASS"
    p = _sass(psk)
    assert "STRONG.SYS" in p            # system-scope peer loads/stores of the gather / broadcast
    ptx_like = subprocess.run(["strings", str(psk)], stdout=subprocess.PIPE, text=True).stdout
    assert "HMMA" not in s.replace("UTCHMMA", "")      # no legacy mma.sync tensor path in the GEMM


def test_sass_of_the_fused_stem_and_gemm_epilogues():
    """The fused stem and the cta_group::2 GEMM (TMA-store epilogue) are real tcgen05 / TMA code, and spill nothing."""
    stem, gexp = ext.OBJ / "stem_kernels.o", ext.OBJ / "bcast_gemm2.o"
    if not stem.exists() or not gexp.exists():
        pytest.skip("object files not present (built elsewhere)")
    s = _sass(stem)
    for mnemonic in ("UTCHMMA", "UTMALDG.2D", "UTMASTG.2D", "LDTM", "LDGSTS"):
        assert mnemonic in s, f"{mnemonic} missing from stem_kernels SASS"
    g = _sass(gexp)
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTMASTG.2D", "LDTM"):
        assert mnemonic in g, f"{mnemonic} missing from bcast_gemm_exp SASS"
    for log in ("stem_kernels.nvcc.log", "bcast_gemm2.nvcc.log", "bcast_gemm.nvcc.log", "bn_kernels.nvcc.log", "pool_kernels.nvcc.log"):
        p = ext.OBJ / log
        if p.exists():
            for line in p.read_text().splitlines():
                if "spill" in line:
                    assert "0 bytes spill stores, 0 bytes spill loads" in line, f"{log}: {line.strip()}"


def test_ps_kernels_spill_budget():
    """``ps_kernels.cu`` at ``__launch_bounds__(256, 3)`` (85 registers): nothing spills except the fp32-wire Adam instantiation
    of the update kernel (four fp32 state vectors + two 16-byte gathers per rank in flight), and that one stays under 128 bytes."""
    p = ext.OBJ / "ps_kernels.nvcc.log"
    if not p.exists():
        pytest.skip("build log not present (built elsewhere)")
    entry, seen = None, 0
    for line in p.read_text().splitlines():
        if "Compiling entry function" in line:
            entry = line.split("'")[1]
        if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
            seen += 1
            assert "psb_update_kernelILi0ELi0ELi1E" in entry, f"{entry}: {line.strip()}"
            assert int(line.split("bytes stack frame,")[1].split("bytes spill stores")[0]) <= 128, line
    assert seen <= 1


def test_built_extensions_are_not_older_than_their_sources():
    """The in-tree ``.so`` files travel to the GPU box as they are: a kernel edited after the last build would be tested (and
    benchmarked) in its OLD form there.  ``make build`` / ``__graft_entry__.build()`` refreshes them."""
    if not ext.CUDA_SO.exists() or not ext.HOST_SO.exists():
        pytest.skip("extensions not built yet (run __graft_entry__.build())")
    native = ext.cuda_sources() + list((ext.CSRC / "kernels").glob("*.cuh")) + list((ext.CSRC / "kernels").glob("*.h")) + \
        list((ext.CSRC / "runtime").glob("*.h")) + [ext.CSRC / "runtime" / "symm_mem.cpp", ext.CSRC / "bindings.cpp",
                                                     ext.CSRC / "gemm_bindings.cpp"]
    def stale():
        return [p.name for p in native if p.stat().st_mtime > ext.CUDA_SO.stat().st_mtime]

    if stale() and shutil.which("g++") and shutil.which(ext.nvcc_path()):
        ext.build_cuda()                  # what `make test` does first (a checkout may also have touched the sources)
    assert not stale(), f"{ext.CUDA_SO.name} is older than {stale()}: rebuild"
    if (ext.CSRC / "runtime" / "host_ext.cpp").stat().st_mtime > ext.HOST_SO.stat().st_mtime and shutil.which("g++"):
        ext.build_host()
    assert (ext.CSRC / "runtime" / "host_ext.cpp").stat().st_mtime <= ext.HOST_SO.stat().st_mtime, "host extension is stale"
