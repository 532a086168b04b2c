// conv_chain.cuh -- a CHAIN of convolutions of the small-map stages (20x20 and 40x40: the C3k blocks, SPPF / C2PSA 1x1
// layers, the P4 / P5 head branches; graph chains 054-190, 195-232, 267-399 of SURVEY.md Appendix A) as ONE launch.
//
// Frames are independent, so nothing in these stages needs a grid-wide dependency: a CTA owns a frame and walks the whole
// chain for it, layer after layer -- TMA load (its own previous output, still in L2) -> tcgen05.mma -> epilogue -> global
// store -> CTA barrier -> next layer.  As separate launches every one of these layers paid a full grid launch / drain
// (4-6 us inside the captured graph, 10-13 us alone) for ~1 us of work on all 148 SMs; measured with the layers left out
// of the graph (tools/probe_skip.py) the two stages cost 0.71 ms of a 1.93 ms step.  A chain keeps 64 CTAs (one per frame
// of the batch) busy and leaves the other SMs to whatever else is in flight (the other head branches, the other runners).
//
// The per-layer machinery is that of conv_tma.cuh (same plans, same swizzled operand layouts, same weight images, same
// issue code): warp 0 = TMA producer, warps 1-4 = MMA issuers (one per 128-row sub-tile), the rest = epilogue.  Per layer
// the CTA re-reads its ConvParams from global memory into shared memory, re-initialises the mbarrier ring and fetches the
// layer's weights; TMEM is allocated once.  Visibility of a layer's output to the next layer's TMA loads: the storing
// threads execute __threadfence() + fence.proxy.async before the CTA barrier that ends the layer.
#pragma once

#include "conv_tma.cuh"

namespace xrseg {

struct alignas(128) ChainLayer {
  TmapSet tmaps;      // input tensor maps of the layer (m[0]; m[0..3] = parity planes in s2 mode)
  ConvParams p;       // plan made for ONE frame per CTA (plan_chain_layer); in / out / res / wpack / bias filled in
};

// Plans layer `d` (d.B = frames of the chunk) for the chain kernel: work items never straddle frames and are as large as
// the TMEM / shared-memory budget allows (one CTA walks all items of a frame).  Returns false when no TMA plan fits.
static inline bool plan_chain_layer(const ConvDesc& d, ConvParams& p) {
  ConvDesc one = d;
  one.B = 1;
  bool ok = false;
  tma_budget_ref() = CONV_SMEM_MAX - 12288;       // room for the kernel's tail region (plans of all layers, second bias buffer)
  if (d.k == 3 && d.stride == 1 && !d.transposed) ok = plan_conv_halo_tma_impl(one, 1, p);
  else if (d.k == 1 && d.stride == 1 && !d.transposed) ok = plan_conv_flat_tma_impl(one, 1, p, /*chain=*/true);
  else if (d.k == 3 && d.stride == 2 && !d.transposed) ok = plan_conv_s2_tma_impl(one, 1, p, false);
  tma_budget_ref() = CONV_SMEM_MAX;
  if (!ok || !p.sw || p.n_tiles > 2 || 2 * p.nsub * p.Ntile > 512) return false;
  p.B = d.B;
  if (p.mode == MODE_FLAT_TMA) {
    p.frame_rows = d.H * d.W;
    p.tpi = ceil_div(p.frame_rows, p.slots);       // items per frame
    p.flat_rows = d.B * p.frame_rows;
  }
  p.M_total = d.B * p.tpi;
  p.m_tiles = p.M_total;
  p.fd_hp1 = make_fastdiv(p.tpi);
  p.grid = d.B;
  return true;
}

#ifdef __CUDACC__
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory header of the chain kernel (inside the CONV_HDR_BYTES every plan leaves free): two mbarrier sets used by
// alternating layers, so the set of layer k+1 is initialised while layer k runs.
struct ChainBars {
  uint64_t full[CONV_MAX_STAGES], empty[CONV_MAX_STAGES], tfull[4], tempty[4], bres;
};
enum { CHAIN_MAX_LAYERS = 20, CHAIN_PARAM_BYTES = (sizeof(ConvParams) + 15) & ~15,
       CHAIN_TAIL_BYTES = CHAIN_MAX_LAYERS * CHAIN_PARAM_BYTES + 2048 };   // plans of all layers + the second bias buffer
static_assert(2 * sizeof(ChainBars) + 16 <= 512 && 512 + 2048 <= CONV_HDR_BYTES + 768, "chain header layout");

__device__ __forceinline__ void chain_init_bars(ChainBars* bs, int n_issuers, bool inval) {
  if (inval) {
    for (int i = 0; i < CONV_MAX_STAGES; ++i) { mbar_inval(&bs->full[i]); mbar_inval(&bs->empty[i]); }
    for (int i = 0; i < 4; ++i) { mbar_inval(&bs->tfull[i]); mbar_inval(&bs->tempty[i]); }
    mbar_inval(&bs->bres);
  }
  for (int i = 0; i < CONV_MAX_STAGES; ++i) {
    mbar_init(&bs->full[i], 1);
    mbar_init(&bs->empty[i], n_issuers);       // every issuer commits the stage it has consumed
  }
  for (int i = 0; i < 4; ++i) {
    mbar_init(&bs->tfull[i], n_issuers);       // ... and the accumulators it has finished
    mbar_init(&bs->tempty[i], TMA_EPI_WARPS);  // one arrival per epilogue warp
  }
  mbar_init(&bs->bres, 1);
  mbar_fence_init();
}
__device__ __forceinline__ void chain_fetch_weights(const ConvParams& p, uint8_t* smem, uint64_t* bres) {
  const uint32_t bytes = static_cast<uint32_t>(p.nks) * p.b_stage_bytes;
  const uint32_t b_dst = smem_u32(smem + p.smem_off_b);
  mbar_arrive_expect_tx(bres, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    bulk_copy_g2s(b_dst + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, n, bres);
  }
}

// Per layer the critical path is: CTA barrier -> TMA load of the layer's input -> MMAs -> epilogue -> fence.  Everything
// else is taken off it: the plans of all layers sit in shared memory from the start, and while layer k runs the idle lanes
// of the producer warp initialise the barrier set and stage the bias of layer k+1, and an epilogue thread starts the fetch
// of layer k+1's resident weights as soon as layer k's last MMA has completed.
__global__ void __launch_bounds__(TMA_THREADS, 1)
conv_chain_kernel(const ChainLayer* __restrict__ layers, int n_layers, int n_frames, int tail_off, long long* dbg) {
  // dbg (libxrseg_debug.so, XRSEG_CHAIN_PROBE=1): clock64() of CTA 0 at five points of every layer, [layer][8]
  extern __shared__ __align__(1024) uint8_t smem[];
  ChainBars* bars = reinterpret_cast<ChainBars*>(smem);                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * sizeof(ChainBars));
  volatile int* prep_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);  // layer counter whose barrier set / bias are ready
  float* bias_buf[2] = {reinterpret_cast<float*>(smem + 512), reinterpret_cast<float*>(smem + tail_off + CHAIN_MAX_LAYERS * CHAIN_PARAM_BYTES)};
  auto plan = [&](int l) -> const ConvParams& {
    return *reinterpret_cast<const ConvParams*>(smem + tail_off + l * CHAIN_PARAM_BYTES);
  };

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;

  // plans of every layer -> shared memory (constants of the launch: no dependence on earlier kernels)
  for (int l = 0; l < n_layers; ++l)
    for (int i = tid; i < static_cast<int>(sizeof(ConvParams) / 4); i += TMA_THREADS)
      reinterpret_cast<uint32_t*>(smem + tail_off + l * CHAIN_PARAM_BYTES)[i] = reinterpret_cast<const uint32_t*>(&layers[l].p)[i];
  if (tid < n_layers) {
    prefetch_tensormap(&layers[tid].tmaps.m[0]);
    if (layers[tid].p.mode == MODE_S2_TMA)
      for (int i = 1; i < 4; ++i) prefetch_tensormap(&layers[tid].tmaps.m[i]);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  if (tid == 0) *prep_flag = -1;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // preparation of the very first layer: barrier set 0, bias, resident weights
  {
    const ConvParams& p0 = plan(0);
    if (tid == 0) {
      chain_init_bars(&bars[0], p0.nsub, false);
      if (p0.b_resident) chain_fetch_weights(p0, smem, &bars[0].bres);
    }
    for (int i = tid; i < p0.n_tiles * p0.Ntile; i += TMA_THREADS) bias_buf[0][i] = p0.act ? 0.5f * p0.bias[i] : p0.bias[i];
  }
  pdl_launch_dependents();
  pdl_wait();                 // everything below reads what earlier kernels wrote

  const int total_layers = ((n_frames - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)) * n_layers;
  int lc = 0;                 // running layer counter of this CTA (over its frames): barrier set / bias buffer = lc & 1
  for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
    for (int l = 0; l < n_layers; ++l, ++lc) {
      // (A) every role is done with the previous layer (its barrier set, shared-memory stages and TMEM columns are free,
      //     its output is in global memory and fenced towards the async proxy); this layer's set and bias are ready
      __syncthreads();
      if (dbg && blockIdx.x == 0 && tid == 0) dbg[lc * 8 + 0] = clock64();
      const ConvParams& p = plan(l);
      const ChainLayer& L = layers[l];
      ChainBars& B = bars[lc & 1];
      uint64_t *full = B.full, *empty = B.empty, *tfull = B.tfull, *tempty = B.tempty, *bres = &B.bres;
      const float* bias_s = bias_buf[lc & 1];
      const int ln = l + 1 < n_layers ? l + 1 : 0;          // the layer that follows (next frame: the chain starts over)
      const bool has_next = lc + 1 < total_layers;
      const int n_issuers = p.nsub;
      const int nbuf = 2;
      uint8_t* smem_b = smem + p.smem_off_b;
      uint8_t* smem_a = smem + p.smem_off_a;
      const int ipf = p.tpi * p.n_tiles;                  // work items of one frame
      const int w0 = f * ipf, w1 = w0 + ipf;

      if (warp == 0) {
        if (lane == 0) {
          // ======================================= TMA producer ========================================
          const uint32_t a_u32 = smem_u32(smem_a);
          const uint32_t b_u32 = smem_u32(smem_b);
          const uint32_t a_tx = static_cast<uint32_t>(p.cps) * p.slots * 16u * (p.kps > 1 ? p.kps : 1) *
                                (p.mode == MODE_S2_TMA ? 4u : 1u);
          const uint32_t b_stride = static_cast<uint32_t>((p.b_stage_bytes + 1023) & ~1023);
          const bool flat = p.mode == MODE_FLAT_TMA;
          int it = 0;
          for (int w = w0; w < w1; ++w) {
            const int tile = p.n_tiles == 1 ? w : (w >> 1);
            const int n_tile = p.n_tiles == 1 ? 0 : (w & 1);
            const int b = fd_div(p.fd_hp1, tile);
            const int t = tile - b * p.tpi;
            const int y0 = t * p.R;
            for (int ks = 0; ks < p.nks; ++ks, ++it) {
              const int slot = it % p.S;
              if (it >= p.S) mbar_wait(&empty[slot], static_cast<uint32_t>((it / p.S) - 1) & 1u);
              mbar_arrive_expect_tx(&full[slot], a_tx + (p.b_resident ? 0u : static_cast<uint32_t>(p.b_stage_bytes)));
              const uint32_t a_dst = a_u32 + slot * p.a_stage_bytes;
              if (p.mode == MODE_S2_TMA) {
                for (int pl = 0; pl < 4; ++pl)
                  tma_load_4d(a_dst + pl * p.lbo_a, &L.tmaps.m[pl], &full[slot], ks * p.cb, -1, y0 - (pl >> 1), b);
              } else if (flat) {
                // rows [b * frame_rows + t * slots, + slots) of the activation matrix: items never straddle frames (rows past
                // the frame belong to the next frame or are zero fill; the epilogue masks them)
                const uint32_t box_bytes = static_cast<uint32_t>(p.hbox) * p.cb * 2u;
                const int row0 = b * p.frame_rows + t * p.slots;
                uint32_t dst = a_dst;
                for (int kb = 0; kb < p.kps; ++kb)
                  for (int r0 = 0; r0 < p.slots; r0 += p.hbox, dst += box_bytes)
                    tma_load_4d(dst, &L.tmaps.m[0], &full[slot], (ks * p.kps + kb) * p.cb, row0 + r0, 0, 0);
              } else {
                tma_load_4d(a_dst, &L.tmaps.m[0], &full[slot], ks * p.cb, -1, y0 - 1, b);
              }
              if (!p.b_resident) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) +
                                     (static_cast<size_t>(n_tile) * p.nks + ks) * p.b_stage_bytes;
                const uint32_t b_dst = b_u32 + slot * b_stride;
                for (uint32_t off = 0; off < static_cast<uint32_t>(p.b_stage_bytes); off += 32768u) {
                  const uint32_t n = p.b_stage_bytes - off < 32768u ? p.b_stage_bytes - off : 32768u;
                  bulk_copy_g2s(b_dst + off, src + off, n, &full[slot]);
                }
              }
            }
          }
        } else if (has_next) {
          // ============ idle lanes of the producer warp: prepare the NEXT layer while this one runs ============
          // (its barrier set was last used two layers ago; its bias buffer too)
          const ConvParams& pn = plan(ln);
          float* bn = bias_buf[(lc + 1) & 1];
          for (int i = lane - 1; i < pn.n_tiles * pn.Ntile; i += 31) bn[i] = pn.act ? 0.5f * pn.bias[i] : pn.bias[i];
          if (lane == 1) {
            chain_init_bars(&bars[(lc + 1) & 1], pn.nsub, lc >= 1);
            __threadfence_block();
            *prep_flag = lc + 1;               // the epilogue thread that prefetches the next weights waits for this
          }
        }
        __syncwarp();   // lane 0 (producer) and lanes 1-31 (preparation) reconverge: bar.sync below needs the whole warp
      } else if (warp < TMA_FIRST_EPI_WARP) {
        // ======================================= MMA issuers =========================================
        const int my_u = warp - 1;
        if (my_u < n_issuers) {
          const uint32_t a_u32 = smem_u32(smem_a);
          const uint32_t b_u32 = smem_u32(smem_b);
          const int kj = p.cb >> 4;
          const uint32_t idesc = p.idesc;
          const uint32_t ntile_u = static_cast<uint32_t>(p.Ntile);
          const uint32_t rb = static_cast<uint32_t>(p.cb) * 2u;
          const uint32_t row16 = rb >> 4;
          const uint32_t sub16 = 128u * row16;
          const uint32_t tap16 = (ntile_u * rb) >> 4;
          const uint32_t b_stride_sw = static_cast<uint32_t>((p.b_stage_bytes + 1023) & ~1023);
          const uint32_t layout = p.sw == 3 ? 2u : (p.sw == 2 ? 4u : 6u);
          const uint64_t hi_sw = static_cast<uint64_t>(((8u * rb) >> 4) | (1u << 14) | (layout << 29)) << 32;
          if (p.b_resident) mbar_wait(bres, 0);
          if (dbg && blockIdx.x == 0 && my_u == 0 && lane == 0) dbg[lc * 8 + 5] = clock64();
          int it = 0, tcount = 0;
          for (int w = w0; w < w1; ++w, ++tcount) {
            const int buf = tcount % nbuf;
            const int use = tcount / nbuf;
            if (use >= 1) mbar_wait(&tempty[buf], static_cast<uint32_t>(use - 1) & 1u);
            tc_fence_after();
            const uint32_t d_base = tmem_base + static_cast<uint32_t>(buf * p.nsub * p.Ntile);
            for (int ks = 0; ks < p.nks; ++ks, ++it) {
              const int slot = it % p.S;
              mbar_wait(&full[slot], static_cast<uint32_t>(it / p.S) & 1u);
              tc_fence_after();
              if (dbg && blockIdx.x == 0 && it == 0 && my_u == 0 && lane == 0) dbg[lc * 8 + 1] = clock64();
              const uint32_t a_base = a_u32 + slot * p.a_stage_bytes;
              const uint32_t b_base = b_u32 + (p.b_resident ? ks * p.b_stage_bytes : slot * b_stride_sw);
              const uint32_t a_lo_stage = ((a_base >> 4) & 0x3FFFu) | (1u << 16);
              const uint32_t b_lo = ((b_base >> 4) & 0x3FFFu) | (1u << 16);
              if (elect_one()) {
                const uint32_t a_sub = a_lo_stage + static_cast<uint32_t>(my_u) * sub16;
                const uint32_t d_tmem = d_base + static_cast<uint32_t>(my_u) * ntile_u;
                const uint32_t acc0 = ks > 0 ? 1u : 0u;
                switch ((p.mode == MODE_HALO_TMA ? 0 : p.mode == MODE_S2_TMA ? 4 : 8) + (kj == 4 ? 2 : kj == 2 ? 1 : 0)) {
                  case 0: issue_taps<0, 1>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 1: issue_taps<0, 2>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 2: issue_taps<0, 4>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 4: issue_taps<1, 1>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 5: issue_taps<1, 2>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 6: issue_taps<1, 4>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 8: issue_taps<2, 1>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  case 9: issue_taps<2, 2>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                  default: issue_taps<2, 4>(d_tmem, a_sub, b_lo, hi_sw, idesc, acc0, row16, tap16, static_cast<uint32_t>(p.Wp) * row16, 2u * (static_cast<uint32_t>(p.lbo_a) >> 4), static_cast<uint32_t>(p.slots) * row16, p.kps > 1 ? p.kps : 1, hi_sw, static_cast<uint32_t>(p.lbo_a) >> 4); break;
                }
                umma_commit(&empty[slot]);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(&tfull[buf]);
            __syncwarp();
          }
        }
      } else {
        // ======================================= epilogue ============================================
        const int ew = warp - TMA_FIRST_EPI_WARP;
        const int q = warp & 3;
        const int grp = ew >> 2;
        constexpr int G = TMA_EPI_GROUPS;
        const int nsub = p.nsub, Ntile = p.Ntile, nch = Ntile >> 4;
        const bool flat = p.mode == MODE_FLAT_TMA, act = p.act != 0;
        const int Wp = p.Wp, Rr = p.R, Hh = p.H, Ww = p.W, slots = p.slots, frame_rows = p.frame_rows, tpi = p.tpi;
        const int n_tiles = p.n_tiles;
        const int out_pitch = p.out_pitch, out2_pitch = p.out2_pitch, res_pitch = p.res_pitch, split_n = p.split_n;
        __half* const outp = p.out;
        __half* const out2p = p.out2;
        const __half* const resp = p.res;
        const FastDiv fd_wp = p.fd_wp, fd_tpi = p.fd_hp1;
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        int gu0 = 0, gc0 = grp;
        while (gc0 >= nch) { gc0 -= nch; ++gu0; }
        int tcount = 0;
        for (int w = w0; w < w1; ++w, ++tcount) {
          const int tile = n_tiles == 1 ? w : (w >> 1);
          const int n_tile = n_tiles == 1 ? 0 : (w & 1);
          const int b = fd_div(fd_tpi, tile);
          const int t = tile - b * tpi;
          const int y0 = t * Rr;
          const int buf = tcount % nbuf;
          const int use = tcount / nbuf;
          int pix[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pix[u] = -1;
            if (u < nsub) {
              const int j = 128 * u + q * 32 + lane;
              if (flat) {
                const int r = t * slots + j;                  // row inside the frame
                if (r < frame_rows) pix[u] = b * frame_rows + r;
              } else {
                const int yy = fd_div(fd_wp, j);
                const int cc = j - yy * Wp;
                const int y = y0 + yy;
                if (yy < Rr && y < Hh && cc >= 1 && cc <= Ww) pix[u] = (b * Hh + y) * Ww + (cc - 1);
              }
            }
          }
          mbar_wait(&tfull[buf], static_cast<uint32_t>(use) & 1u);
          tc_fence_after();
          if (dbg && blockIdx.x == 0 && w == w1 - 1 && tid == TMA_FIRST_EPI_WARP * 32) dbg[lc * 8 + 2] = clock64();
          if (w == w1 - 1 && has_next && tid == TMA_FIRST_EPI_WARP * 32) {
            // every MMA of this layer has completed: its weights (and operand stages) are dead.  Start the fetch of the next
            // layer's resident weights now, under this layer's epilogue and fence.
            const ConvParams& pn = plan(ln);
            if (pn.b_resident) {
              while (*prep_flag < lc + 1) { }      // the next barrier set is initialised (done at this layer's start)
              __threadfence_block();
              chain_fetch_weights(pn, smem, &bars[(lc + 1) & 1].bres);
            }
          }
          __syncwarp();
          if (dbg && blockIdx.x == 0 && w == w1 - 1 && tid == TMA_FIRST_EPI_WARP * 32) dbg[lc * 8 + 6] = clock64();
          const uint32_t acc = lane_taddr + static_cast<uint32_t>(buf * nsub * Ntile);
          auto finish_unit = [&](const uint32_t (&v)[16], int u, int c) {
            const int px = u == 0 ? pix[0] : (u == 1 ? pix[1] : (u == 2 ? pix[2] : pix[3]));
            if (px < 0) return;
            const int n = n_tile * Ntile + c * 16;
            __half* dst;
            if (split_n && n >= split_n) dst = out2p + static_cast<size_t>(px) * out2_pitch + (n - split_n);
            else dst = outp + static_cast<size_t>(px) * out_pitch + n;
            epilogue_chunk16_hb(v, bias_s + n, act, resp ? resp + static_cast<size_t>(px) * res_pitch + n : nullptr, dst);
          };
          auto advance = [&](int& u, int& c) {
            c += G;
            while (c >= nch) { c -= nch; ++u; }
          };
          int u = gu0, c = gc0;
          uint32_t va[16], vb[16];
          if (u < nsub) tmem_ld16(acc + static_cast<uint32_t>(u * Ntile + c * 16), va);
          while (u < nsub) {
            int u2 = u, c2 = c;
            advance(u2, c2);
            tmem_ld_wait();
            if (u2 < nsub) tmem_ld16(acc + static_cast<uint32_t>(u2 * Ntile + c2 * 16), vb);
            if (dbg && blockIdx.x == 0 && w == w1 - 1 && tid == TMA_FIRST_EPI_WARP * 32 && u == gu0 && c == gc0) dbg[lc * 8 + 7] = clock64();
            finish_unit(va, u, c);
            if (u2 >= nsub) break;
            u = u2; c = c2;
            advance(u, c);
            tmem_ld_wait();
            if (u < nsub) tmem_ld16(acc + static_cast<uint32_t>(u * Ntile + c * 16), va);
            finish_unit(vb, u2, c2);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        // this layer's output -> visible to the next layer's TMA loads (async proxy) and residual reads of this CTA
        if (dbg && blockIdx.x == 0 && tid == TMA_FIRST_EPI_WARP * 32) dbg[lc * 8 + 3] = clock64();
        __threadfence();
        fence_proxy_async_all();
        if (dbg && blockIdx.x == 0 && tid == TMA_FIRST_EPI_WARP * 32) dbg[lc * 8 + 4] = clock64();
      }
      tc_fence_before();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static inline void conv_chain_prepare_device() {
  XR_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_MAX));
}

// smem_bytes: max over the chain's layers of p.smem_bytes; the plans of all layers and the second bias buffer live behind it
static inline void launch_conv_chain(const ChainLayer* d_layers, int n_layers, int n_frames, int smem_bytes, int num_sms,
                                     cudaStream_t stream, long long* dbg = nullptr) {
  const int tail_off = (smem_bytes + 15) & ~15;
  const int total = tail_off + CHAIN_TAIL_BYTES;
  XR_CHECK(n_layers <= CHAIN_MAX_LAYERS && total <= CONV_SMEM_MAX, "chain of %d layers needs %d bytes of shared memory", n_layers, total);
  launch_k(conv_chain_kernel, n_frames < num_sms ? n_frames : num_sms, TMA_THREADS, static_cast<size_t>(total), stream, d_layers,
           n_layers, n_frames, tail_off, dbg);
}
#endif  // __CUDACC__

}  // namespace xrseg
