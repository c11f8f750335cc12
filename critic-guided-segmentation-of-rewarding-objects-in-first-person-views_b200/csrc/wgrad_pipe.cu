// Persistent, cp.async double-buffered weight-gradient kernel for the encoder convolutions
// (features.0/3/6 of NewCritic at chfak 1: Cin <= 8, Cout == 8, dY = ReLU+MaxPool backward of a pooled gradient).
//
// Same maths as wgrad_mma.cu (mma.sync m16n8k8 TF32, A fragments gathered from the haloed input tile, a row of
// ones for the bias gradient) but restructured around the memory system, because ncu showed the staged version
// spending its life in `long_scoreboard` (load -> sync -> MMA -> reduce, ~0.1 of HBM peak):
//   * operands travel HBM -> shared memory as RAW bytes with 16-byte cp.async (zero register staging, zero index
//     maths per element): full-width image rows of X (pixel-major, with zeroed left/right pads standing in for the
//     horizontal halo, zfill for rows outside the image) and the pooled dE / E / argmax rows;
//   * the ReLU + max-pool backward is evaluated while fetching the B fragment
//     (dY[y,x,co] = (argmax == window position && E > 0) ? dE : 0), so the full-resolution gradient never exists;
//   * a 4-stage ring per CTA: the copies of tiles i+1..i+3 are in flight while the warps run the MMAs of tile i;
//   * ONE persistent CTA per SM; accumulators live in registers across all of a CTA's tiles and are reduced once at
//     the end.  This matters more than anything else here: every CTA adds into the SAME few hundred weight-gradient
//     addresses and same-address REDs serialise in L2 at ~50 ns each, so the staged kernels' run time was
//     (#CTAs x 50 ns) -- 1024 / 512 / 256 CTAs -> 46 / 27 / 14.6 us -- whatever the rest of the kernel did.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cgs {

struct WpGeom {
  int th, tiles_y, ntiles;
  int pitch, padL, rowsX;        // X smem row pitch (floats), left pad (floats), rows per tile (th + 2)
  int xchunks, pchunks, ichunks; // 16-byte chunks per X row, per pooled dE/E row, per pooled idx row
  int prow;                      // floats per pooled row (W/2 * 8)
  int stage_floats;              // floats per pipeline stage
  int offE, offI;                // float offsets of E and idx inside a stage (dE follows X)
  int offD;
  int flags;                     // debug: bit 0 = skip the final REDs (tools/wgrad_probe.py)
  int u8, raw_row, raw_chunks, offRaw, rollc;   // X given as raw uint8 frames (+ roll): raw rows staged, converted per tile
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t to_tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ void mma_tf32_16n8k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MT>
__global__ void __launch_bounds__(256) wgrad3x3_pipe_kernel(const cgs_wgrad3x3_args p, const WpGeom g) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int H = p.H, W = p.W, C = p.x.C, H2 = H >> 1, W2 = W >> 1;
  const int th = g.th;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

  constexpr int NS = 3;
  float* s_lut = smem + NS * g.stage_floats;      // (float)i / 255.0f, i = 0..255 (raw uint8 frames only)
  if (g.u8) s_lut[tid] = (float)tid / 255.0f;
  // zero the horizontal-halo pads of all stages once (cp.async never touches them)
  for (int e = tid; e < NS * g.rowsX * 2 * g.padL; e += 256) {
    const int s = e / (g.rowsX * 2 * g.padL), r2 = e - s * (g.rowsX * 2 * g.padL);
    const int r = r2 / (2 * g.padL), q = r2 - r * (2 * g.padL);
    const int col = q < g.padL ? q : g.padL + W * C + (q - g.padL);
    smem[s * g.stage_floats + r * g.pitch + col] = 0.f;
  }

  auto prefetch = [&](int tile, int s) {
    const int n = tile / g.tiles_y, y0 = (tile - n * g.tiles_y) * th;
    const uint32_t sb = smem_base + (uint32_t)(s * g.stage_floats) * 4u;
    if (g.u8) {
      // raw uint8 rows (W*3 bytes each); converted to fp32 (/255, roll) by a shared->shared pass once they have landed
      const uint8_t* fr = reinterpret_cast<const uint8_t*>(p.x.a);
      for (int c = tid; c < g.rowsX * g.raw_chunks; c += 256) {
        const int r = c / g.raw_chunks, cx = c - r * g.raw_chunks;
        const int y = y0 - 1 + r;
        const bool ok = y >= 0 && y < H;
        const uint8_t* src = ok ? fr + (size_t)(unsigned)((n * H + y) * W) * 3u + cx * 16 : fr;
        cp_async16(sb + (uint32_t)g.offRaw * 4u + (uint32_t)(r * g.raw_row + cx * 16), src, ok ? 16 : 0);
      }
    } else
    // X: rowsX full-width image rows (row r <-> image row y0 - 1 + r), pixel-major
    for (int c = tid; c < g.rowsX * g.xchunks; c += 256) {
      const int r = c / g.xchunks, cx = c - r * g.xchunks;
      const int y = y0 - 1 + r;
      const bool ok = y >= 0 && y < H;
      const float* src = ok ? p.x.a + (size_t)(unsigned)((n * H + y) * W) * (unsigned)C + cx * 4 : p.x.a;
      cp_async16(sb + (uint32_t)(r * g.pitch + g.padL) * 4u + (uint32_t)cx * 16u, src, ok ? 16 : 0);
    }
    // pooled dE, E rows (8 channels per pooled pixel) and the argmax bytes
    const int pr = th >> 1;
    const size_t pbase = (size_t)(unsigned)((n * H2 + (y0 >> 1)) * W2) * 8u;
    for (int c = tid; c < pr * g.pchunks; c += 256) {
      cp_async16(sb + (uint32_t)g.offD * 4u + (uint32_t)c * 16u, p.dy.a + pbase + (size_t)c * 4, 16);
      cp_async16(sb + (uint32_t)g.offE * 4u + (uint32_t)c * 16u, p.dy.b + pbase + (size_t)c * 4, 16);
    }
    for (int c = tid; c < pr * g.ichunks; c += 256)
      cp_async16(sb + (uint32_t)g.offI * 4u + (uint32_t)c * 16u, p.dy.idx + pbase + (size_t)c * 16, 16);
    cp_async_commit();
  };

  // per-lane A rows, tap-major: m = tap*C + ci  (8 consecutive rows = 8 channels of one tap: conflict-free gathers)
  // Branch-free inner loop: value = x[off] * mulA + addA with (mulA, addA) = (1,0) data row, (0,1) ones row, (0,0) pad row
  int offA[MT][2];
  float mulA[MT][2], addA[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = 16 * mt + gid + 8 * h;
      int off = g.padL;
      float mu = 0.f, ad = 0.f;
      if (m < 9 * C) {
        const int tap = m / C, ci = m - tap * C;
        off = (tap / 3) * g.pitch + g.padL + ((tap % 3) - 1) * C + ci;
        mu = 1.f;
      } else if (m == 9 * C) {
        ad = 1.f;
      }
      offA[mt][h] = off; mulA[mt][h] = mu; addA[mt][h] = ad;
    }
  float acc[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[mt][j] = 0.f;

  for (int k = 0; k < NS - 1; ++k) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t < g.ntiles) prefetch(t, k); else cp_async_commit();      // empty groups keep the accounting uniform
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
    const int s = it % NS;
    const int ahead = tile + (NS - 1) * gridDim.x;
    if (ahead < g.ntiles) prefetch(ahead, (it + NS - 1) % NS); else cp_async_commit();
    cp_async_wait<NS - 1>();        // everything but the newest NS-1 groups has landed: this tile is in smem
    __syncthreads();
    if (g.u8) {
      // raw bytes -> fp32 X rows.  Pixel-major fp32 mirrors the byte layout, so this is a flat expansion with a rotation by
      // roll*3 bytes; (float)u8/255.0f comes from a 256-entry table (exact, no divisions), 4 bytes -> one 128-bit store.
      int roll = g.rollc;
      if (p.x.b) { roll = *reinterpret_cast<const int*>(p.x.b) % W; if (roll < 0) roll += W; }
      const int rb = roll * 3, rowb = g.raw_row, q4 = rowb >> 2;
      float* dX = smem + s * g.stage_floats;
      const uint8_t* raw = reinterpret_cast<const uint8_t*>(dX + g.offRaw);
      for (int e = tid; e < g.rowsX * q4; e += 256) {
        const int r = e / q4, i = (e - r * q4) * 4;
        const uint8_t* rr = raw + r * rowb;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int sb = i + k + rb;
          if (sb >= rowb) sb -= rowb;
          v[k] = s_lut[rr[sb]];
        }
        *reinterpret_cast<float4*>(dX + r * g.pitch + g.padL + i) = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncthreads();
    }
    const float* sX = smem + s * g.stage_floats;
    const float* sD = sX + g.offD;
    const float* sE = sX + g.offE;
    const unsigned char* sI = reinterpret_cast<const unsigned char*>(sX + g.offI);
    for (int yl = warp; yl < th; yl += 8) {
      const float* xrow = sX + yl * g.pitch + tig * C;
      const int prow_off = (yl >> 1) * g.prow + gid;
      const int posy = (yl & 1) << 1;
      for (int xb = 0; xb < W; xb += 8) {
        // B fragment: dY[y, xb + tig (+4), co = gid] through ReLU + max-pool backward
        uint32_t b[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int x = xb + tig + 4 * h;
          const int pe = prow_off + (x >> 1) * 8;
          float v = sD[pe];
          v = (sI[pe] == (posy | (x & 1))) ? v : 0.f;      // selects, not branches: the loop must stay convergent
          v = (sE[pe] > 0.f) ? v : 0.f;
          b[h] = to_tf32(v);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int off = offA[mt][h];
            const float v0 = fmaf(xrow[off + xb * C], mulA[mt][h], addA[mt][h]);
            const float v1 = fmaf(xrow[off + (xb + 4) * C], mulA[mt][h], addA[mt][h]);
            a[h] = to_tf32(v0);
            a[2 + h] = to_tf32(v1);
          }
          mma_tf32_16n8k8(acc[mt], a[0], a[1], a[2], a[3], b[0], b[1]);
        }
      }
    }
    __syncthreads();   // everyone is done with stage s before the next iteration's prefetch overwrites it
  }

  // ---- reduction: 8 warps -> this CTA's smem -> (thread-block cluster, distributed shared memory) -> one RED per weight
  //      per CLUSTER.  Every CTA adds into the same few hundred addresses and same-address REDs serialise in L2 at
  //      ~45 ns each (tools/wgrad_probe.py: +6.5 us at 148 CTAs, +25 us at 592), so the cluster leader sums its peers'
  //      partial tiles over DSMEM first: 4x fewer RED rounds.
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  float* s_red = smem;                    // [8 warps][16*MT][8]
  float* s_tot = smem + 8 * 16 * MT * 8;  // [16*MT][8] this CTA's total
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float* d = s_red + (warp * 16 * MT + 16 * mt + gid) * 8 + 2 * tig;
    d[0] = acc[mt][0]; d[1] = acc[mt][1];
    d[64] = acc[mt][2]; d[65] = acc[mt][3];
  }
  __syncthreads();
  for (int e = tid; e < 16 * MT * 8; e += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w * 16 * MT * 8 + e];
    s_tot[e] = sum;
  }
  cluster.sync();
  if (cluster.block_rank() == 0) {
    const unsigned nb = cluster.num_blocks();
    for (int e = tid; e < 16 * MT * 8; e += 256) {
      const int m = e >> 3, co = e & 7;
      float sum = s_tot[e];
      for (unsigned r = 1; r < nb; ++r) sum += cluster.map_shared_rank(s_tot, r)[e];
      if (g.flags & 1) {
        if (sum == 123.456f) p.dw[0] = sum;     // debug probe: keep the reduction alive without the REDs
      } else if (m < 9 * C) {
        const int tap = m / C, ci = m - tap * C;
        atomicAdd(p.dw + ((size_t)co * C + ci) * 9 + tap, sum);
      } else if (m == 9 * C && p.db) {
        atomicAdd(p.db + co, sum);
      }
    }
  }
  cluster.sync();   // peers keep their shared memory alive until the leader has read it
}

bool wgrad_pipe_supported(const cgs_wgrad3x3_args& a) {
  static const int nopipe = getenv("CGS_WGRAD_NOPIPE") ? atoi(getenv("CGS_WGRAD_NOPIPE")) : 0;   // debug probe
  if (nopipe) return false;
  const bool u8 = a.x.mode == CGS_SRC_U8ROLL && a.x.C == 3 && (a.W % 16) == 0;
  if (!u8 && (a.x.mode != CGS_SRC_PLAIN || a.x.b != nullptr)) return false;
  if (a.dy.mode != CGS_SRC_POOLBWD) return false;
  // measured (tools/wgrad_probe.py, B=256): 34.8 vs 37.9 us on the RGB layer, but 24.6/18.4 vs 22.5/16.4 us on the 8->8
  // layers, where the staged persistent kernel's 32-pixel-wide planar tiles need fewer shared-memory wavefronts per MMA
  if (a.dy.C != 8 || a.x.C > 4 || a.x.C < 1) return false;
  if (a.H < 16 || a.W < 16 || (a.H % 16) || (a.W % 8)) return false;
  if ((a.W * a.x.C) % 4) return false;                          // X rows must be whole 16-byte chunks
  if (((a.W / 2) * 8) % 16) return false;                       // argmax rows too
  if ((reinterpret_cast<uintptr_t>(a.x.a) | reinterpret_cast<uintptr_t>(a.dy.a) | reinterpret_cast<uintptr_t>(a.dy.b) |
       reinterpret_cast<uintptr_t>(a.dy.idx)) & 15) return false;
  return true;
}

int launch_wgrad_pipe(const cgs_wgrad3x3_args& a, cudaStream_t st) {
  WpGeom g;
  const int C = a.x.C, W = a.W;
  g.th = 16;
  g.tiles_y = a.H / g.th;
  g.ntiles = a.B * g.tiles_y;
  g.padL = (C + 3) & ~3;
  g.pitch = g.padL + W * C + g.padL;
  g.pitch = (g.pitch + 3) & ~3;
  if ((g.pitch & 31) == 0) g.pitch += 4;                        // rows of a tap triple should not alias banks
  g.rowsX = g.th + 2;
  g.xchunks = W * C / 4;
  g.prow = (W / 2) * 8;
  g.pchunks = g.prow / 4;
  g.ichunks = g.prow / 16;
  const int pr = g.th / 2;
  g.offD = g.rowsX * g.pitch;
  g.offE = g.offD + pr * g.prow;
  g.offI = g.offE + pr * g.prow;
  g.offRaw = g.offI + (pr * g.prow + 3) / 4;
  g.offRaw = (g.offRaw + 3) & ~3;
  g.u8 = a.x.mode == CGS_SRC_U8ROLL;
  g.raw_row = W * 3;
  g.raw_chunks = g.raw_row / 16;
  g.rollc = ((a.x.shift % W) + W) % W;
  g.stage_floats = g.offRaw + (g.u8 ? (g.rowsX * g.raw_row + 3) / 4 : 0);
  g.stage_floats = (g.stage_floats + 3) & ~3;
  size_t smem = (size_t)3 * g.stage_floats * sizeof(float) + 256 * sizeof(float);
  const int MT = (9 * C + 1 + 15) / 16;
  const size_t red = (size_t)9 * 16 * (MT <= 2 ? 2 : 5) * 8 * sizeof(float);
  if (smem < red) smem = red;
  cudaFuncSetAttribute(wgrad3x3_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(wgrad3x3_pipe_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sms = device_sms();
  int grid = sms;   // one persistent CTA per SM (the non-RED part of the kernel does not depend on the grid size)
  g.flags = 0;
  if (const char* e = getenv("CGS_WGRAD_NORED")) g.flags |= atoi(e) ? 1 : 0;      // debug probes only
  if (const char* e = getenv("CGS_WGRAD_GRID")) { const int v = atoi(e); if (v > 0) grid = v; }
  if (grid > g.ntiles) grid = g.ntiles;
  // thread-block clusters of 4 (148 = 4 x 37): the grid is rounded down to a multiple of the cluster size
  int csz = 2;
  if (const char* e = getenv("CGS_WGRAD_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) csz = v; }
  if (grid < csz) csz = 1;
  grid -= grid % csz;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = MT <= 2 ? cudaLaunchKernelEx(&cfg, wgrad3x3_pipe_kernel<2>, a, g)
                            : cudaLaunchKernelEx(&cfg, wgrad3x3_pipe_kernel<5>, a, g);
  if (err != cudaSuccess) { set_error("wgrad3x3_pipe: launch failed: %s", cudaGetErrorString(err)); return CGS_ECUDA; }
  return check_launch("wgrad3x3_pipe");
}

}  // namespace cgs
