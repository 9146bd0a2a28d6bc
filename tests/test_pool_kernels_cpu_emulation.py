"""Run the REAL source of ``psb_maxpool_fwd_rows`` / ``psb_maxpool_bwd_quads`` (``csrc/kernels/pool_kernels.cu``) on the CPU.

Both kernels are barrier-free and shared-memory-free, so executing their bodies once per (blockIdx, threadIdx) in a plain C++
loop is an exact emulation.  The test cuts the two kernels out of the ``.cu`` file, compiles them with g++ against a 40-line
shim (``uint4``, ``__nv_bfloat16`` as raw bits, ``pack_bf16x2`` / ``unpack_bf16x8`` with round-to-nearest-even) and compares
the result with ``F.max_pool2d`` + autograd — the kernels were written after the round's GPU budget was spent, so this is
their numerics check until the GPU suite (``tests/test_gpu_bn.py::test_maxpool_matches_torch``) runs."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F


@pytest.fixture(autouse=True)
def _single_threaded_torch():
    """The emulated kernels run on the calling thread; keeping torch single-threaded makes the timings of these tests repeatable."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pytorch_ps_mpi_b200", "csrc", "kernels", "pool_kernels.cu")

SHIM = r'''
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
typedef uint16_t __nv_bfloat16;
struct Idx { int x; };
static Idx blockIdx, threadIdx, gridDim, blockDim;
#define __global__
#define __launch_bounds__(...)
#define __restrict__
static inline void unpack_bf16x8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  for (int i = 0; i < 4; ++i) {
    uint32_t lo = w[i] << 16, hi = w[i] & 0xffff0000u;
    memcpy(&f[2 * i], &lo, 4);
    memcpy(&f[2 * i + 1], &hi, 4);
  }
}
static inline uint32_t bf16_bits(float a) {           // round to nearest even, like __floats2bfloat162_rn
  uint32_t u; memcpy(&u, &a, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0u;
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
static inline uint32_t pack_bf16x2(float a, float b) { return bf16_bits(a) | (bf16_bits(b) << 16); }
struct PoolGeom { int N, H, W, C, OH, OW, groups; };
'''

MAIN = r'''
int main(int argc, char** argv) {
  // argv: N C H W mode(0 fwd, 1 bwd) grid in out
  PoolGeom g; g.N = atoi(argv[1]); g.C = atoi(argv[2]); g.H = atoi(argv[3]); g.W = atoi(argv[4]);
  g.OH = (g.H + 2 - 3) / 2 + 1; g.OW = (g.W + 2 - 3) / 2 + 1; g.groups = g.C / 8;
  const int mode = atoi(argv[5]), grid = atoi(argv[6]);
  const int lanes = 256 / g.groups;
  const size_t nin = (size_t)g.N * g.H * g.W * g.C, nout = (size_t)g.N * g.OH * g.OW * g.C;
  FILE* fi = fopen(argv[7], "rb"); FILE* fo = fopen(argv[8], "wb");
  gridDim.x = grid; blockDim.x = 256;
  if (mode == 0) {
    std::vector<uint16_t> x(nin), y(nout); std::vector<uint8_t> arg(nout);
    if (fread(x.data(), 2, nin, fi) != nin) return 2;
    for (int b = 0; b < grid; ++b) for (int t = 0; t < 256; ++t) {
      blockIdx.x = b; threadIdx.x = t;
      psb_maxpool_fwd_rows(x.data(), y.data(), arg.data(), g, lanes);
    }
    fwrite(y.data(), 2, nout, fo); fwrite(arg.data(), 1, nout, fo);
  } else {
    std::vector<uint16_t> dy(nout), dx(nin, 0x7fc0); std::vector<uint8_t> arg(nout);
    if (fread(dy.data(), 2, nout, fi) != nout || fread(arg.data(), 1, nout, fi) != nout) return 2;
    for (int b = 0; b < grid; ++b) for (int t = 0; t < 256; ++t) {
      blockIdx.x = b; threadIdx.x = t;
      psb_maxpool_bwd_quads(dy.data(), arg.data(), dx.data(), g, lanes);
    }
    fwrite(dx.data(), 2, nin, fo);
  }
  fclose(fi); fclose(fo);
  return 0;
}
'''


def _cut(src: str, name: str) -> str:
    a = src.index("__global__ void __launch_bounds__(256) " + name)
    depth, i = 0, src.index("{", a)
    while True:
        depth += src[i] == "{"
        depth -= src[i] == "}"
        i += 1
        if depth == 0:
            return src[a:i]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = open(SRC).read()
    body = _cut(src, "psb_maxpool_fwd_rows") + "\n" + _cut(src, "psb_maxpool_bwd_quads")
    d = tmp_path_factory.mktemp("poolemu")
    (d / "emu.cpp").write_text(SHIM + body + MAIN)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(d / "emu"), str(d / "emu.cpp")], check=True)
    return d


def _bf16(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.float32)).bfloat16()


@pytest.mark.parametrize("shape,grid", [((2, 64, 12, 16), 7), ((1, 16, 6, 4), 3), ((3, 8, 2, 2), 1), ((2, 32, 8, 10), 50)])
def test_real_pool_kernel_source_matches_torch(emu, shape, grid):
    N, C, H, W = shape
    rng = np.random.default_rng(0)
    x = _bf16(rng.standard_normal((N, C, H, W)))                                  # NCHW logical
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    fin, fout = emu / "in.bin", emu / "out.bin"
    fin.write_bytes(x_nhwc.view(torch.int16).numpy().tobytes())
    subprocess.run([str(emu / "emu"), str(N), str(C), str(H), str(W), "0", str(grid), str(fin), str(fout)], check=True)
    raw = fout.read_bytes()
    nout = N * OH * OW * C
    y = torch.from_numpy(np.frombuffer(raw[: 2 * nout], dtype=np.int16).copy()).view(torch.bfloat16).view(N, OH, OW, C)
    arg = np.frombuffer(raw[2 * nout:], dtype=np.uint8).copy()
    xt = x.float().requires_grad_(True)
    yt = F.max_pool2d(xt, 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2).float(), yt.detach())
    if H % 2 == 0 and W % 2 == 0:
        dy = _bf16(rng.standard_normal((N, C, OH, OW)))
        fin.write_bytes(dy.permute(0, 2, 3, 1).contiguous().view(torch.int16).numpy().tobytes() + arg.tobytes())
        subprocess.run([str(emu / "emu"), str(N), str(C), str(H), str(W), "1", str(grid), str(fin), str(fout)], check=True)
        dx = torch.from_numpy(np.frombuffer(fout.read_bytes(), dtype=np.int16).copy()).view(torch.bfloat16).view(N, H, W, C)
        yt.backward(dy.float())
        want = xt.grad.bfloat16().float()          # overlapping windows can send two gradients to one pixel: fp32 sum → bf16
        assert torch.equal(dx.permute(0, 3, 1, 2).float(), want)
