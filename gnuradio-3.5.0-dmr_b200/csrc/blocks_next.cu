// The callers either side of the hot path (SURVEY.md 8f rank 4 and VERDICT round 1 "missing" 1, 2, 6):
//   gr_framer_sink_1              (gnuradio-core/src/lib/general/gr_framer_sink_1.cc:90-178, .h:60-103)
//   digital_clock_recovery_mm_cc  (gr-digital/lib/digital_clock_recovery_mm_cc.cc:117-213, .h:75-80)
//   gr_map_bb, gr_unpack_k_bits_bb, gr_stream_to_streams, gr_vector_to_streams as blocks of their own
//                                 (general/gr_map_bb.cc:35-61, gr_unpack_k_bits_bb.cc:38-70,
//                                  gr_stream_to_streams.cc:37-66, gr_vector_to_streams.cc:37-70)
// All are byte / index / feedback-loop work: bit exact against the oracle, no tensor cores, HBM or latency bound.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gr_math.cuh"
#include "internal.h"

using namespace grb;

namespace {

struct Plan {
  cudaStream_t stream = nullptr;
  Stager stager;
  DevBuf d_in, d_out;
  std::mutex mu;
  int init() {
    GRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return GRCUDA_OK;
  }
  cudaStream_t pick(void* s) const { return s ? (cudaStream_t)s : stream; }
  virtual ~Plan() { if (stream) cudaStreamDestroy(stream); }
};

bool have_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error(GRCUDA_ECUDA, "no CUDA device available (%s); libgr_cuda has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return false;
  }
  return true;
}

int grid_of(long items, int threads, int per_sm = 8) {
  const long g = (items + threads - 1) / threads;
  return (int)std::max<long>(1, std::min(g, (long)sm_count() * per_sm));
}

// =====================================================================================================================
// gr_map_bb: out[i] = d_map[in[i]]  (gr_map_bb.cc:49-61).  2 B of HBM per item; 16 items per thread and access.
// =====================================================================================================================
__global__ void __launch_bounds__(256) map_bb_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                     long n, const unsigned char* __restrict__ map) {
  __shared__ unsigned char tab[256];
  tab[threadIdx.x] = map[threadIdx.x];
  __syncthreads();
  const long nvec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) ? 0 : n / 16;
  const uint4* in4 = reinterpret_cast<const uint4*>(in);
  uint4* out4 = reinterpret_cast<uint4*>(out);
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    uint4 x = __ldg(in4 + v);
    unsigned w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; q++)
      w[q] = tab[w[q] & 255] | (tab[(w[q] >> 8) & 255] << 8) | (tab[(w[q] >> 16) & 255] << 16) | (tab[w[q] >> 24] << 24);
    out4[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (long i = nvec * 16 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = tab[in[i]];
}

// =====================================================================================================================
// gr_unpack_k_bits_bb: byte i -> k bytes, most significant of the k low bits first (gr_unpack_k_bits_bb.cc:53-70).
// A thread makes 4 consecutive output bytes (one 32-bit store); (1 + 1/k) B of HBM per output item.
// =====================================================================================================================
// k in {1, 2, 4, 8}: a thread expands 16 / k input bytes into one 16-byte store
template <int K>
__global__ void __launch_bounds__(256) unpack_pow2_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                          long nvec) {
  constexpr int NIN = 16 / K;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    unsigned char t[NIN];
    if (NIN == 16) *reinterpret_cast<uint4*>(t) = __ldg(reinterpret_cast<const uint4*>(in) + v);
    else if (NIN == 8) *reinterpret_cast<uint2*>(t) = __ldg(reinterpret_cast<const uint2*>(in) + v);
    else if (NIN == 4) *reinterpret_cast<unsigned*>(t) = __ldg(reinterpret_cast<const unsigned*>(in) + v);
    else *reinterpret_cast<unsigned short*>(t) = __ldg(reinterpret_cast<const unsigned short*>(in) + v);
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int o = 0; o < 16; o++) {
      const unsigned bit = (t[o / K] >> (K - 1 - (o % K))) & 1u;
      w[o >> 2] |= bit << (8 * (o & 3));
    }
    reinterpret_cast<uint4*>(out)[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(256) unpack_k_bits_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                            long nout, unsigned k) {
  const bool al = (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  const long nvec = al ? nout / 4 : 0;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    unsigned w = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long o = v * 4 + q;
      const long i = o / k;
      const unsigned j = k - 1 - (unsigned)(o - i * k);
      const unsigned t = __ldg(in + i);
      w |= ((j < 32 ? (t >> j) : 0u) & 1u) << (8 * q);
    }
    reinterpret_cast<unsigned*>(out)[v] = w;
  }
  for (long o = nvec * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x; o < nout; o += (long)gridDim.x * blockDim.x) {
    const long i = o / k;
    const unsigned j = k - 1 - (unsigned)(o - i * k);
    const unsigned t = in[i];
    out[o] = (unsigned char)((j < 32 ? (t >> j) : 0u) & 1u);
  }
}

// =====================================================================================================================
// gr_stream_to_streams / gr_vector_to_streams: item i of stream j = in[i * nstreams + j]
// (gr_stream_to_streams.cc:52-66, gr_vector_to_streams.cc:53-70: the same loop).  A 32 x 32 item tile through shared
// memory so that both the read (along j) and the write (along i) are contiguous; 2 x item_size B of HBM per item.
// =====================================================================================================================
template <class T>
__global__ void __launch_bounds__(256) deinterleave_kernel(const T* __restrict__ in, T* __restrict__ out, long nitems,
                                                           int nstreams, long out_stride) {
  __shared__ T tile[32][33];
  const long tiles_i = (nitems + 31) / 32;
  const int tiles_j = (nstreams + 31) / 32;
  const long ntiles = tiles_i * tiles_j;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long i0 = (t / tiles_j) * 32;
    const int j0 = (int)(t % tiles_j) * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const long i = i0 + r;
      const int j = j0 + tx;
      if (i < nitems && j < nstreams) tile[r][tx] = in[i * nstreams + j];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r;
      const long i = i0 + tx;
      if (i < nitems && j < nstreams) out[(long)j * out_stride + i] = tile[tx][r];
    }
    __syncthreads();
  }
}
// any other item size: byte granular
__global__ void __launch_bounds__(256) deinterleave_bytes_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                                 long nitems, int nstreams, int item_size, long out_stride) {
  const long total = nitems * nstreams * item_size;
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long)gridDim.x * blockDim.x) {
    const long item = id / item_size;
    const int b = (int)(id - item * item_size);
    const long i = item / nstreams;
    const int j = (int)(item - i * nstreams);
    out[((long)j * out_stride + i) * item_size + b] = in[id];
  }
}

// =====================================================================================================================
// gr_framer_sink_1 (gr_framer_sink_1.cc:90-178): sync flag -> 32-bit header (two identical 16-bit halves: 4 bits
// whitener offset, 12 bits payload length) -> payload bytes MSB first -> message.  One state machine per channel.
// =====================================================================================================================
struct FramerChan {
  int state;               // 0 SYNC_SEARCH, 1 HAVE_SYNC, 2 HAVE_HEADER
  unsigned header;
  int headerbitlen_cnt;
  int packetlen, whitener_offset, packetlen_cnt, byte_index;
  unsigned packet_byte;
  long long consumed;      // bytes of this channel's stream seen so far
  int seq, pad;            // messages posted by this channel
};
struct FramerQueue {       // device-side gr_msg_queue: records + payload arena, appended with atomics
  int nmsgs, dropped;
  unsigned long long payload_used;
};
struct FramerArgs {
  const unsigned char* in;   // item t of channel c = in[t * t_stride + c * c_stride]
  long t_stride, c_stride;
  long nitems;               // per channel, unless counts
  const int* counts;         // [nchan] or nullptr
  int count_scale;           // valid items of channel c = counts[c] * count_scale
  int nchan;
  FramerChan* chan;
  unsigned char* packets;    // [nchan][4096] payload being assembled (MAX_PKT_LEN, .h:62)
  FramerQueue* q;
  grcuda_framer_msg* msgs;
  int max_msgs;
  unsigned char* payload;
  unsigned long long payload_cap;
};

// post one message (gr_make_message(0, whitener_offset, 0, len) + insert_tail, :139-146 and :170-178)
__device__ __forceinline__ void framer_post(const FramerArgs& a, int c, FramerChan& s, const unsigned char* pkt, int len,
                                            long long end_index) {
  const int slot = atomicAdd(&a.q->nmsgs, 1);
  const int seq = s.seq++;
  if (slot >= a.max_msgs) { atomicAdd(&a.q->dropped, 1); return; }
  grcuda_framer_msg m;
  m.channel = c; m.whitener_offset = s.whitener_offset; m.length = len; m.seq = seq;
  m.end_index = end_index;
  m.payload_offset = 0;
  if (len > 0) {
    const unsigned long long off = atomicAdd(&a.q->payload_used, (unsigned long long)len);
    if (off + (unsigned long long)len > a.payload_cap) {
      atomicAdd(&a.q->payload_used, (unsigned long long)(-(long long)len));   // give the reservation back
      m.payload_offset = -1;           // arena full: the record survives, the bytes do not
      atomicAdd(&a.q->dropped, 1);
    } else {
      m.payload_offset = (long long)off;
      for (int i = 0; i < len; i++) a.payload[off + i] = pkt[i];
    }
  }
  a.msgs[slot] = m;
}

// one step of the reference's loop for one input byte; returns nothing, state in s
__device__ __forceinline__ void framer_byte(const FramerArgs& a, int c, FramerChan& s, unsigned char* pkt, unsigned b,
                                            long long index) {
  if (s.state == 0) {                       // :104-112 -- the flagged byte is NOT consumed by the search: it is the
    if (!(b & 2)) return;                   // first header bit
    s.state = 1; s.header = 0; s.headerbitlen_cnt = 0;   // enter_have_sync (:46-55)
  }
  if (s.state == 1) {                       // :121-159
    s.header = (s.header << 1) | (b & 1);
    if (++s.headerbitlen_cnt == 32) {
      if (((s.header >> 16) ^ (s.header & 0xffff)) == 0) {     // header_ok (.h:88-92)
        s.packetlen = (int)((s.header >> 16) & 0x0fff);        // header_payload (.h:94-102)
        s.whitener_offset = (int)((s.header >> 28) & 0x000f);
        s.state = 2; s.packetlen_cnt = 0; s.packet_byte = 0; s.byte_index = 0;   // enter_have_header (:57-69)
        if (s.packetlen == 0) { framer_post(a, c, s, pkt, 0, index); s.state = 0; }
      } else {
        s.state = 0;
      }
    }
    return;
  }
  s.packet_byte = ((s.packet_byte << 1) | (b & 1)) & 0xff;     // :161-187
  if (s.byte_index++ == 7) {
    pkt[s.packetlen_cnt++] = (unsigned char)s.packet_byte;
    s.byte_index = 0;
    if (s.packetlen_cnt == s.packetlen) { framer_post(a, c, s, pkt, s.packetlen_cnt, index); s.state = 0; }
  }
}

// [time][channel] layout (what the batched correlator writes): one thread per channel, a warp reads 32 neighbouring
// bytes of a row; 16 rows are loaded before they are stepped through so that the loads overlap.
__global__ void __launch_bounds__(128) framer_rows_kernel(const FramerArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nchan) return;
  FramerChan s = a.chan[c];
  unsigned char* pkt = a.packets + (size_t)c * 4096;
  const long n = a.counts ? (long)a.counts[c] * a.count_scale : a.nitems;
  const unsigned char* p = a.in + (long)c * a.c_stride;
  long t = 0;
  // groups of 16 rows, the NEXT group's loads in flight while this one is looked at: with 250 warps on the machine a
  // group is one DRAM round trip, and nothing else hides it (a warp-load is one 32-byte sector of a row)
  unsigned nx[16];
  if (n >= 16) {
#pragma unroll
    for (int i = 0; i < 16; i++) nx[i] = __ldg(p + (long)i * a.t_stride);
  }
  for (; t + 16 <= n; t += 16) {
    unsigned b[16];
#pragma unroll
    for (int i = 0; i < 16; i++) b[i] = nx[i];
    if (t + 32 <= n) {
#pragma unroll
      for (int i = 0; i < 16; i++) nx[i] = __ldg(p + (t + 16 + i) * a.t_stride);
    }
    if (s.state == 0) {                     // searching and no flag in these 16 items: nothing happens (:104-112)
      unsigned any = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) any |= b[i];
      if (!(any & 2)) continue;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) framer_byte(a, c, s, pkt, b[i], s.consumed + t + i);
  }
  for (; t < n; t++) framer_byte(a, c, s, pkt, __ldg(p + t * a.t_stride), s.consumed + t);
  s.consumed += n;
  a.chan[c] = s;
}

// stream-major layout (t_stride == 1; the reference's single stream is nchan == 1): one WARP per channel.  The search
// for the next flag runs 512 bytes per step (16 per lane, one ballot); payload bytes are packed 32 at a time (lane l
// turns items [8l, 8l+8) into one byte) whenever the byte counter is aligned; lane 0 steps through the rest.
__global__ void __launch_bounds__(128) framer_stream_kernel(const FramerArgs a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= a.nchan) return;
  FramerChan s = a.chan[c];                  // every lane keeps a copy; lane 0's is the one that advances and is shared
  unsigned char* pkt = a.packets + (size_t)c * 4096;
  const long n = a.counts ? (long)a.counts[c] * a.count_scale : a.nitems;
  const unsigned char* p = a.in + (long)c * a.c_stride;
  long t = 0;
  while (t < n) {
    if (s.state == 0) {
      // skip to the first byte with the flag bit (lanes look at 16 bytes each)
      bool found = false;
      while (t < n && !found) {
        const long base = t + lane * 16;
        unsigned m = 0;
        for (int i = 0; i < 16; i++)
          if (base + i < n && (__ldg(p + base + i) & 2)) { m = (unsigned)i + 1; break; }
        const unsigned who = __ballot_sync(0xffffffffu, m != 0);
        if (who) {
          const int l = __ffs(who) - 1;
          t = t + l * 16 + (long)__shfl_sync(0xffffffffu, m, l) - 1;
          found = true;
        } else {
          t = t + 512 < n ? t + 512 : n;
        }
      }
      if (!found) break;
    } else if (s.state == 2 && s.byte_index == 0 && n - t >= 8 && s.packetlen_cnt + 1 < s.packetlen) {
      // whole payload bytes that do NOT complete the packet: 8 items -> 1 byte per lane
      const int want = (int)min((long)min(32, s.packetlen - 1 - s.packetlen_cnt), (n - t) / 8);
      if (lane < want) {
        unsigned v = 0;
        for (int i = 0; i < 8; i++) v = (v << 1) | (__ldg(p + t + lane * 8 + i) & 1u);
        pkt[s.packetlen_cnt + lane] = (unsigned char)v;
      }
      s.packetlen_cnt += want;
      t += (long)want * 8;
      __syncwarp();
      continue;
    }
    // header bits, the last payload byte, ragged ends: the reference's loop, one byte at a time, on lane 0
    long adv = 0;
    if (lane == 0) {
      while (t + adv < n) {
        framer_byte(a, c, s, pkt, __ldg(p + t + adv), s.consumed + t + adv);
        adv++;
        if (s.state == 0 || (s.state == 2 && s.byte_index == 0 && s.packetlen_cnt + 1 < s.packetlen)) break;
      }
    }
    __syncwarp();
    adv = __shfl_sync(0xffffffffu, adv, 0);
    t += adv;
    s.state = __shfl_sync(0xffffffffu, s.state, 0);
    s.byte_index = __shfl_sync(0xffffffffu, s.byte_index, 0);
    s.packetlen = __shfl_sync(0xffffffffu, s.packetlen, 0);
    s.packetlen_cnt = __shfl_sync(0xffffffffu, s.packetlen_cnt, 0);
  }
  if (lane == 0) {
    s.consumed += n;
    a.chan[c] = s;
  }
}

// =====================================================================================================================
// digital_clock_recovery_mm_cc (digital_clock_recovery_mm_cc.cc:117-213): the complex Mueller & Mueller loop with the
// 8-tap complex MMSE interpolator (gri_mmse_fir_interpolator_cc.cc:61-71 over gr_fir_ccf, generic summation order).
// One thread per channel, [time][channel] layout; a feedback loop on float data, restated operation by operation.
// =====================================================================================================================
struct MMCCChan {
  float mu, omega;
  float2 p_2T, p_1T, p_0T, c_2T, c_1T, c_0T;
  long long next_abs;        // absolute index of the next input item the loop reads
  int clamped, overflow;     // steps that wanted to go back before the buffered rows / calls stopped by max_out
};
struct MMCCArgs {
  const float2* in;          // row 0 has absolute index abs_row0
  long ninput, abs_row0;
  float2* out;               // [max_out][nchan]
  float* err;                // [max_out][nchan] or nullptr
  int max_out, nchan;
  int* counts;
  MMCCChan* chan;
  float gain_omega, gain_mu, omega_mid, omega_relative_limit;
  const float* mmse;         // [129][8], coefficient applied to in[ii + i]
};

// The loop feeds floor(mu) back into the next input index, so a channel is one thread; what the kernel can do is keep
// HBM latency out of the recursion: a CTA owns 64 neighbouring channels and walks the rows in tiles of MMCC_TR rows
// (8 rows of overlap = the interpolator's window) that arrive in shared memory through cp.async one tile ahead of the
// one being consumed.  Inside a tile every thread runs its own recursion on LDS.64 reads until its window would leave
// the tile.  HBM sees each input row once (8 B per item) plus the symbols.
#define MMCC_TR 40
#define MMCC_CH 64
__device__ __forceinline__ void mmcc_load_tile(const MMCCArgs& a, float2* __restrict__ buf, long t0, int c0) {
  const int c = c0 + threadIdx.x;
  if (c >= a.nchan) return;
  const float2* src = a.in + t0 * (long)a.nchan + c;
  for (int r = 0; r < MMCC_TR; r++) {
    if (t0 + r < a.ninput) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + r * MMCC_CH + threadIdx.x);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + (long)r * a.nchan) : "memory");
    }
  }
}

__global__ void __launch_bounds__(MMCC_CH) mm_cc_kernel(const MMCCArgs a) {
  __shared__ __align__(16) float2 tiles[2][MMCC_TR * MMCC_CH];
  __shared__ float tab[129 * 8];
  for (int i = threadIdx.x; i < 129 * 8; i += MMCC_CH) tab[i] = a.mmse[i];
  const int c0 = blockIdx.x * MMCC_CH;
  const int c = c0 + threadIdx.x;
  const bool live = c < a.nchan;
  MMCCChan s;
  if (live) s = a.chan[c];
  long ii = 0;
  if (live) {
    ii = (long)(s.next_abs - a.abs_row0);
    if (ii < 0) { ii = 0; s.clamped++; }
  }
  const long ni = a.ninput - 8 - 16;                 // ntaps() and FUDGE (:124)
  const float lim = a.err ? 4.0f : 1.0f;             // :146 vs :178
  const long M = a.nchan;
  constexpr int S = MMCC_TR - 8;                     // rows a tile is responsible for
  const long ntiles = ni > 0 ? (ni + S - 1) / S : 0;
  int oo = 0;
  if (ntiles > 0) mmcc_load_tile(a, tiles[0], 0, c0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (long k = 0; k < ntiles; k++) {
    if (k + 1 < ntiles) mmcc_load_tile(a, tiles[(k + 1) & 1], (k + 1) * S, c0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float2* __restrict__ tile = tiles[k & 1] + threadIdx.x;
    const long t0 = k * S;
    const long stop = (k + 1) * S < ni ? (k + 1) * S : ni;
    while (live && oo < a.max_out && ii < stop) {
      s.p_2T = s.p_1T;
      s.p_1T = s.p_0T;
      const float* e = tab + mm_imu(s.mu) * 8;
      const float2* x = tile + (ii - t0) * MMCC_CH;
      float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;   // gr_fir_ccf_generic::filter (gr_fir_XXX_generic.cc.t:28-55)
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 v0 = x[i * MMCC_CH], v1 = x[(i + 1) * MMCC_CH];
        const float t0c = e[i], t1c = e[i + 1];
        a0r = GR_FADD(a0r, GR_FMUL(t0c, v0.x)); a0i = GR_FADD(a0i, GR_FMUL(t0c, v0.y));
        a1r = GR_FADD(a1r, GR_FMUL(t1c, v1.x)); a1i = GR_FADD(a1i, GR_FMUL(t1c, v1.y));
      }
      s.p_0T = make_float2(GR_FADD(a0r, a1r), GR_FADD(a0i, a1i));
      s.c_2T = s.c_1T;
      s.c_1T = s.c_0T;
      s.c_0T = make_float2(s.p_0T.x > 0 ? 1.0f : 0.0f, s.p_0T.y > 0 ? 1.0f : 0.0f);   // slicer_0deg (:93-103)
      // x = (c_0T - c_2T) * conj(p_1T); y = (p_0T - p_2T) * conj(c_1T); mm_val = real(y - x)   (:137-140)
      const float ar = GR_FSUB(s.c_0T.x, s.c_2T.x), ai = GR_FSUB(s.c_0T.y, s.c_2T.y);
      const float xr = GR_FSUB(GR_FMUL(ar, s.p_1T.x), GR_FMUL(ai, -s.p_1T.y));
      const float br = GR_FSUB(s.p_0T.x, s.p_2T.x), bi = GR_FSUB(s.p_0T.y, s.p_2T.y);
      const float yr = GR_FSUB(GR_FMUL(br, s.c_1T.x), GR_FMUL(bi, -s.c_1T.y));
      float mm_val = GR_FSUB(yr, xr);
      a.out[(long)oo * M + c] = s.p_0T;
      mm_val = branchless_clip(mm_val, lim);
      s.omega = GR_FADD(s.omega, GR_FMUL(a.gain_omega, mm_val));
      s.omega = GR_FADD(a.omega_mid, branchless_clip(GR_FSUB(s.omega, a.omega_mid), a.omega_relative_limit));
      s.mu = GR_FADD(GR_FADD(s.mu, s.omega), GR_FMUL(a.gain_mu, mm_val));
      const float fl = floorf(s.mu);
      ii += (long)(int)fl;
      s.mu = GR_FSUB(s.mu, fl);
      if (a.err) a.err[(long)oo * M + c] = mm_val;
      oo++;
      if (ii < t0) {                                   // a step BACK out of the tile: only bogus input does that (:166)
        if (ii < 0) { ii = 0; s.clamped++; }
        if (ii < t0) { ii = t0; s.clamped++; }         // (the tile's first row is the oldest one still on chip)
      }
    }
    __syncthreads();                                   // tile k's buffer is refilled with tile k + 2 next
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!live) return;
  if (oo >= a.max_out && ii < ni) s.overflow++;
  s.next_abs = a.abs_row0 + ii;
  a.counts[c] = oo;
  a.chan[c] = s;
}

}  // namespace

// =====================================================================================================================
// plans + C ABI
// =====================================================================================================================
struct grcuda_map_bb : Plan {
  DevBuf d_map;
};
struct grcuda_unpack_k_bits : Plan {
  unsigned k = 1;
};
struct grcuda_streams : Plan {   // stream_to_streams and vector_to_streams: the same data movement
  std::vector<unsigned char> h_tmp;
  size_t item_size = 1;
  int nstreams = 1;
  int launch(const void* d_in, void* d_out, long nitems, long out_stride, cudaStream_t s) {
    if (nitems <= 0) return GRCUDA_OK;
    const long tiles = ((nitems + 31) / 32) * ((nstreams + 31) / 32);
    const int g = (int)std::max<long>(1, std::min<long>(tiles, (long)sm_count() * 16));
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out) | (uintptr_t)(out_stride * item_size);
    if (item_size == 16 && !(al & 15)) deinterleave_kernel<uint4><<<g, 256, 0, s>>>((const uint4*)d_in, (uint4*)d_out, nitems, nstreams, out_stride);
    else if (item_size == 8 && !(al & 7)) deinterleave_kernel<uint2><<<g, 256, 0, s>>>((const uint2*)d_in, (uint2*)d_out, nitems, nstreams, out_stride);
    else if (item_size == 4 && !(al & 3)) deinterleave_kernel<unsigned><<<g, 256, 0, s>>>((const unsigned*)d_in, (unsigned*)d_out, nitems, nstreams, out_stride);
    else if (item_size == 2 && !(al & 1)) deinterleave_kernel<unsigned short><<<g, 256, 0, s>>>((const unsigned short*)d_in, (unsigned short*)d_out, nitems, nstreams, out_stride);
    else if (item_size == 1) deinterleave_kernel<unsigned char><<<g, 256, 0, s>>>((const unsigned char*)d_in, (unsigned char*)d_out, nitems, nstreams, out_stride);
    else deinterleave_bytes_kernel<<<grid_of(nitems * nstreams * (long)item_size, 256), 256, 0, s>>>((const unsigned char*)d_in, (unsigned char*)d_out, nitems, nstreams, (int)item_size, out_stride);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};
struct grcuda_framer : Plan {
  int nchan = 1, max_msgs = 0;
  size_t payload_cap = 0;
  DevBuf d_chan, d_packets, d_q, d_msgs, d_payload;
  int launch(const unsigned char* d_in, long nitems, long t_stride, long c_stride, const int* d_counts, int count_scale, cudaStream_t s) {
    FramerArgs a;
    a.in = d_in; a.t_stride = t_stride; a.c_stride = c_stride; a.nitems = nitems; a.counts = d_counts; a.count_scale = count_scale;
    a.nchan = nchan; a.chan = d_chan.as<FramerChan>(); a.packets = d_packets.as<unsigned char>(); a.q = d_q.as<FramerQueue>();
    a.msgs = d_msgs.as<grcuda_framer_msg>(); a.max_msgs = max_msgs; a.payload = d_payload.as<unsigned char>(); a.payload_cap = payload_cap;
    if (t_stride == 1) framer_stream_kernel<<<(nchan * 32 + 127) / 128, 128, 0, s>>>(a);
    else framer_rows_kernel<<<(nchan + 127) / 128, 128, 0, s>>>(a);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};
struct grcuda_mm_cc : Plan {
  int nchan = 1;
  float omega0 = 0, gain_omega = 0, mu0 = 0, gain_mu = 0, lim = 0, omega_mid = 0;
  DevBuf d_chan, d_counts, d_err;
  DeviceTables tables;
  void set_omega_host(float omega) {   // set_omega (.h:75-80): double expressions stored to float members
    const float mn = (float)(omega * (1.0 - lim)), mx = (float)(omega * (1.0 + lim));
    omega_mid = (float)(0.5 * (mn + mx));
  }
  int launch(const float2* d_in, long ninput, long abs_row0, float2* d_out, float* d_err, int max_out, int* d_cnt, cudaStream_t s) {
    MMCCArgs a;
    a.in = d_in; a.ninput = ninput; a.abs_row0 = abs_row0; a.out = d_out; a.err = d_err; a.max_out = max_out; a.nchan = nchan;
    a.counts = d_cnt; a.chan = d_chan.as<MMCCChan>(); a.gain_omega = gain_omega; a.gain_mu = gain_mu; a.omega_mid = omega_mid;
    a.omega_relative_limit = lim; a.mmse = tables.mmse_eff;
    mm_cc_kernel<<<(nchan + MMCC_CH - 1) / MMCC_CH, MMCC_CH, 0, s>>>(a);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

extern "C" {

// ---- gr_map_bb ------------------------------------------------------------------------------------------------------
grcuda_map_bb* grcuda_map_bb_create(const int* map, int nmap) {
  if (nmap < 0 || (nmap > 0 && !map)) { set_error(GRCUDA_EINVAL, "map_bb: bad map"); return nullptr; }
  if (!have_device()) return nullptr;
  unsigned char tab[256];
  for (int i = 0; i < 256; i++) tab[i] = (unsigned char)i;                       // :40-41
  for (int i = 0; i < std::min(nmap, 256); i++) tab[i] = (unsigned char)map[i];  // :43-45
  grcuda_map_bb* h = new grcuda_map_bb;
  if (h->init() || h->d_map.reserve(256) || cudaMemcpy(h->d_map.p, tab, 256, cudaMemcpyHostToDevice) != cudaSuccess) { delete h; return nullptr; }
  return h;
}
void grcuda_map_bb_destroy(grcuda_map_bb* h) { delete h; }
int grcuda_map_bb_work_device(grcuda_map_bb* h, long noutput_items, const unsigned char* d_in, unsigned char* d_out, void* stream) {
  if (noutput_items <= 0) return GRCUDA_OK;
  map_bb_kernel<<<grid_of((noutput_items + 15) / 16, 256), 256, 0, h->pick(stream)>>>(d_in, d_out, noutput_items, h->d_map.as<unsigned char>());
  GRB_LAUNCH_CHECK();
  return GRCUDA_OK;
}
int grcuda_map_bb_work(grcuda_map_bb* h, int noutput_items, const unsigned char* in, unsigned char* out) {
  if (noutput_items <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve(noutput_items)) || (rc = h->d_out.reserve(noutput_items))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, noutput_items, h->stream))) return rc;
  if ((rc = grcuda_map_bb_work_device(h, noutput_items, h->d_in.as<unsigned char>(), h->d_out.as<unsigned char>(), h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, noutput_items, h->stream))) return rc;
  return noutput_items;
}

// ---- gr_unpack_k_bits_bb --------------------------------------------------------------------------------------------
grcuda_unpack_k_bits* grcuda_unpack_k_bits_bb_create(unsigned k) {
  if (k == 0) { set_error(GRCUDA_ERANGE, "interpolation must be > 0"); return nullptr; }   // :44-45 std::out_of_range
  if (!have_device()) return nullptr;
  grcuda_unpack_k_bits* h = new grcuda_unpack_k_bits;
  h->k = k;
  if (h->init()) { delete h; return nullptr; }
  return h;
}
void grcuda_unpack_k_bits_bb_destroy(grcuda_unpack_k_bits* h) { delete h; }
unsigned grcuda_unpack_k_bits_bb_interpolation(grcuda_unpack_k_bits* h) { return h->k; }
int grcuda_unpack_k_bits_bb_work_device(grcuda_unpack_k_bits* h, long noutput_items, const unsigned char* d_in, unsigned char* d_out, void* stream) {
  const long nout = noutput_items / h->k * h->k;     // :60: noutput_items / d_k input bytes
  if (nout <= 0) return GRCUDA_OK;
  cudaStream_t s = h->pick(stream);
  const unsigned k = h->k;
  long done = 0;   // output bytes made by the 16-byte kernel
  if ((k == 1 || k == 2 || k == 4 || k == 8) && !(reinterpret_cast<uintptr_t>(d_out) & 15) && !(reinterpret_cast<uintptr_t>(d_in) & (16 / k - 1)) &&
      nout >= 16) {
    const long nvec = nout / 16;
    const int g = grid_of(nvec, 256);
    if (k == 1) unpack_pow2_kernel<1><<<g, 256, 0, s>>>(d_in, d_out, nvec);
    else if (k == 2) unpack_pow2_kernel<2><<<g, 256, 0, s>>>(d_in, d_out, nvec);
    else if (k == 4) unpack_pow2_kernel<4><<<g, 256, 0, s>>>(d_in, d_out, nvec);
    else unpack_pow2_kernel<8><<<g, 256, 0, s>>>(d_in, d_out, nvec);
    GRB_LAUNCH_CHECK();
    done = nvec * 16;
  }
  if (done < nout) {
    unpack_k_bits_kernel<<<grid_of((nout - done + 3) / 4, 256), 256, 0, s>>>(d_in + done / k, d_out + done, nout - done, k);
    GRB_LAUNCH_CHECK();
  }
  return GRCUDA_OK;
}
int grcuda_unpack_k_bits_bb_work(grcuda_unpack_k_bits* h, int noutput_items, const unsigned char* in, unsigned char* out) {
  if (noutput_items <= 0) return 0;
  const long nin = noutput_items / h->k, nout = nin * h->k;
  if (nout <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve(nin)) || (rc = h->d_out.reserve(nout))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, nin, h->stream))) return rc;
  if ((rc = grcuda_unpack_k_bits_bb_work_device(h, nout, h->d_in.as<unsigned char>(), h->d_out.as<unsigned char>(), h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, nout, h->stream))) return rc;
  return (int)nout;
}

// ---- gr_stream_to_streams / gr_vector_to_streams --------------------------------------------------------------------
static grcuda_streams* streams_new(size_t item_size, size_t nstreams) {
  if (item_size == 0 || nstreams == 0 || nstreams > (1u << 24)) { set_error(GRCUDA_EINVAL, "item_size and nstreams must be > 0"); return nullptr; }
  if (!have_device()) return nullptr;
  grcuda_streams* h = new grcuda_streams;
  h->item_size = item_size;
  h->nstreams = (int)nstreams;
  if (h->init()) { delete h; return nullptr; }
  return h;
}
grcuda_streams* grcuda_stream_to_streams_create(size_t item_size, size_t nstreams) { return streams_new(item_size, nstreams); }
grcuda_streams* grcuda_vector_to_streams_create(size_t item_size, size_t nstreams) { return streams_new(item_size, nstreams); }
void grcuda_streams_destroy(grcuda_streams* h) { delete h; }
int grcuda_streams_nstreams(grcuda_streams* h) { return h->nstreams; }
int grcuda_streams_work_device(grcuda_streams* h, long noutput_items, const void* d_in, void* d_out, long out_stride_items, void* stream) {
  if (out_stride_items < noutput_items) return set_error(GRCUDA_EINVAL, "streams: out_stride_items %ld < noutput_items %ld", out_stride_items, noutput_items);
  return h->launch(d_in, d_out, noutput_items, out_stride_items, h->pick(stream));
}
int grcuda_streams_work(grcuda_streams* h, int noutput_items, const void* in, void* const* out) {
  if (noutput_items <= 0) return 0;
  const size_t total = (size_t)noutput_items * h->nstreams * h->item_size;
  int rc;
  if ((rc = h->d_in.reserve(total)) || (rc = h->d_out.reserve(total))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, total, h->stream))) return rc;
  if ((rc = h->launch(h->d_in.p, h->d_out.p, noutput_items, noutput_items, h->stream))) return rc;
  const size_t per = (size_t)noutput_items * h->item_size;
  if (h->nstreams <= 8) {
    for (int j = 0; j < h->nstreams; j++)
      if ((rc = h->stager.d2h(out[j], (const char*)h->d_out.p + (size_t)j * per, per, h->stream))) return rc;
  } else {   // many streams: ONE device-to-host transfer (each d2h synchronises), then the scatter to the caller's buffers
    h->h_tmp.resize(total);
    if ((rc = h->stager.d2h(h->h_tmp.data(), h->d_out.p, total, h->stream))) return rc;
    for (int j = 0; j < h->nstreams; j++) memcpy(out[j], h->h_tmp.data() + (size_t)j * per, per);
  }
  return noutput_items;
}

// ---- gr_framer_sink_1 -----------------------------------------------------------------------------------------------
grcuda_framer* grcuda_framer_sink_1_create(int nchan, int max_msgs, size_t payload_capacity) {
  if (nchan < 1 || max_msgs < 1) { set_error(GRCUDA_EINVAL, "framer_sink_1: nchan and max_msgs must be >= 1"); return nullptr; }
  if (!have_device()) return nullptr;
  grcuda_framer* h = new grcuda_framer;
  h->nchan = nchan;
  h->max_msgs = max_msgs;
  h->payload_cap = std::max<size_t>(payload_capacity, 16);
  if (h->init() || h->d_chan.reserve((size_t)nchan * sizeof(FramerChan)) || h->d_packets.reserve((size_t)nchan * 4096) ||
      h->d_q.reserve(sizeof(FramerQueue)) || h->d_msgs.reserve((size_t)max_msgs * sizeof(grcuda_framer_msg)) ||
      h->d_payload.reserve(h->payload_cap) ||
      cudaMemset(h->d_chan.p, 0, (size_t)nchan * sizeof(FramerChan)) != cudaSuccess ||   // enter_search() (:86)
      cudaMemset(h->d_q.p, 0, sizeof(FramerQueue)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    delete h;
    return nullptr;
  }
  return h;
}
void grcuda_framer_sink_1_destroy(grcuda_framer* h) { delete h; }
int grcuda_framer_sink_1_work_device(grcuda_framer* h, long nitems, const unsigned char* d_in, long item_stride, long chan_stride,
                                     const int* d_counts, int count_scale, void* stream) {
  if (nitems <= 0 && !d_counts) return GRCUDA_OK;
  if (item_stride < 1) return set_error(GRCUDA_EINVAL, "framer_sink_1: item_stride must be >= 1");
  return h->launch(d_in, nitems, item_stride, chan_stride, d_counts, count_scale, h->pick(stream));
}
int grcuda_framer_sink_1_work(grcuda_framer* h, int noutput_items, const unsigned char* in) {
  if (h->nchan != 1) return set_error(GRCUDA_EINVAL, "host work() is the single-stream form (nchan == 1)");
  if (noutput_items <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve(noutput_items))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, noutput_items, h->stream))) return rc;
  if ((rc = h->launch(h->d_in.as<unsigned char>(), noutput_items, 1, 0, nullptr, 1, h->stream))) return rc;
  GRB_CUDA(cudaStreamSynchronize(h->stream));
  return noutput_items;   // a sync block: consumes everything it is given (:189-190)
}
int grcuda_framer_sink_1_count(grcuda_framer* h, int* dropped) {
  FramerQueue q;
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpy(&q, h->d_q.p, sizeof q, cudaMemcpyDeviceToHost));
  if (dropped) *dropped = q.dropped;
  return std::min(q.nmsgs, h->max_msgs);
}
int grcuda_framer_sink_1_read(grcuda_framer* h, grcuda_framer_msg* msgs, int max_msgs, unsigned char* payload, size_t payload_cap,
                              int* dropped) {
  FramerQueue q;
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpy(&q, h->d_q.p, sizeof q, cudaMemcpyDeviceToHost));
  const int have = std::min(q.nmsgs, h->max_msgs);
  if (dropped) *dropped = q.dropped;
  std::vector<grcuda_framer_msg> all((size_t)have);
  if (have) GRB_CUDA(cudaMemcpy(all.data(), h->d_msgs.p, (size_t)have * sizeof(grcuda_framer_msg), cudaMemcpyDeviceToHost));
  const size_t used = (size_t)std::min<unsigned long long>(q.payload_used, h->payload_cap);
  std::vector<unsigned char> bytes(used);
  if (used) GRB_CUDA(cudaMemcpy(bytes.data(), h->d_payload.p, used, cudaMemcpyDeviceToHost));
  // arrival order of one shared queue: by the stream position that completed the packet, then by channel
  std::sort(all.begin(), all.end(), [](const grcuda_framer_msg& x, const grcuda_framer_msg& y) {
    if (x.end_index != y.end_index) return x.end_index < y.end_index;
    if (x.channel != y.channel) return x.channel < y.channel;
    return x.seq < y.seq;
  });
  if (have > max_msgs) return set_error(GRCUDA_EINVAL, "framer_sink_1: %d messages queued, room for %d", have, max_msgs);
  size_t off = 0;
  for (int i = 0; i < have; i++) {
    grcuda_framer_msg m = all[i];
    if (m.payload_offset >= 0 && m.length > 0) {
      if (off + (size_t)m.length > payload_cap) return set_error(GRCUDA_EINVAL, "framer_sink_1: payload buffer too small");
      memcpy(payload + off, bytes.data() + m.payload_offset, (size_t)m.length);
      m.payload_offset = (long long)off;
      off += (size_t)m.length;
    } else if (m.payload_offset >= 0) {
      m.payload_offset = (long long)off;
    }
    msgs[i] = m;
  }
  // the queue is emptied (delete_head until empty); the per-channel state machines keep going
  FramerQueue z = {0, 0, 0};
  GRB_CUDA(cudaMemcpy(h->d_q.p, &z, sizeof z, cudaMemcpyHostToDevice));
  return have;
}

// ---- digital_clock_recovery_mm_cc -----------------------------------------------------------------------------------
grcuda_mm_cc* grcuda_clock_recovery_mm_cc_create(int nchan, float omega, float gain_omega, float mu, float gain_mu,
                                                 float omega_relative_limit) {
  if (omega <= 0.0f) { set_error(GRCUDA_ERANGE, "clock rate must be > 0"); return nullptr; }                        // :65-66
  if (gain_mu < 0 || gain_omega < 0) { set_error(GRCUDA_ERANGE, "Gains must be non-negative"); return nullptr; }   // :67-68
  if (nchan < 1) { set_error(GRCUDA_EINVAL, "clock_recovery_mm_cc: nchan must be >= 1"); return nullptr; }
  if (!have_device()) return nullptr;
  grcuda_mm_cc* h = new grcuda_mm_cc;
  h->nchan = nchan; h->omega0 = omega; h->gain_omega = gain_omega; h->mu0 = mu; h->gain_mu = gain_mu; h->lim = omega_relative_limit;
  h->set_omega_host(omega);
  std::vector<MMCCChan> st((size_t)nchan);
  for (auto& s : st) { memset(&s, 0, sizeof s); s.mu = mu; s.omega = omega; }
  if (h->init() || get_tables(&h->tables) || h->d_chan.reserve((size_t)nchan * sizeof(MMCCChan)) || h->d_counts.reserve((size_t)nchan * sizeof(int)) ||
      cudaMemcpy(h->d_chan.p, st.data(), (size_t)nchan * sizeof(MMCCChan), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    delete h;
    return nullptr;
  }
  return h;
}
void grcuda_clock_recovery_mm_cc_destroy(grcuda_mm_cc* h) { delete h; }
int grcuda_clock_recovery_mm_cc_forecast(grcuda_mm_cc* h, int noutput_items) {   // :84-91, with the block's CURRENT omega of channel 0
  MMCCChan s;
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpy(&s, h->d_chan.p, sizeof s, cudaMemcpyDeviceToHost));
  return (int)ceil((noutput_items * s.omega) + 8) + 16;
}
int grcuda_clock_recovery_mm_cc_get_state(grcuda_mm_cc* h, int chan, float* mu, float* omega) {
  if (chan < 0 || chan >= h->nchan) return set_error(GRCUDA_EINVAL, "channel %d out of range", chan);
  MMCCChan s;
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpy(&s, (char*)h->d_chan.p + (size_t)chan * sizeof s, sizeof s, cudaMemcpyDeviceToHost));
  if (mu) *mu = s.mu;
  if (omega) *omega = s.omega;
  return GRCUDA_OK;
}
static int mmcc_edit(grcuda_mm_cc* h, bool set_mu, float v) {
  std::lock_guard<std::mutex> lk(h->mu);
  GRB_CUDA(cudaDeviceSynchronize());
  std::vector<MMCCChan> st((size_t)h->nchan);
  GRB_CUDA(cudaMemcpy(st.data(), h->d_chan.p, st.size() * sizeof(MMCCChan), cudaMemcpyDeviceToHost));
  for (auto& s : st) { if (set_mu) s.mu = v; else s.omega = v; }
  GRB_CUDA(cudaMemcpy(h->d_chan.p, st.data(), st.size() * sizeof(MMCCChan), cudaMemcpyHostToDevice));
  if (!set_mu) h->set_omega_host(v);
  return GRCUDA_OK;
}
int grcuda_clock_recovery_mm_cc_set_mu(grcuda_mm_cc* h, float mu) { return mmcc_edit(h, true, mu); }
int grcuda_clock_recovery_mm_cc_set_omega(grcuda_mm_cc* h, float omega) { return mmcc_edit(h, false, omega); }
int grcuda_clock_recovery_mm_cc_set_gain_mu(grcuda_mm_cc* h, float g) { cudaDeviceSynchronize(); h->gain_mu = g; return GRCUDA_OK; }
int grcuda_clock_recovery_mm_cc_set_gain_omega(grcuda_mm_cc* h, float g) { cudaDeviceSynchronize(); h->gain_omega = g; return GRCUDA_OK; }
int grcuda_clock_recovery_mm_cc_counters(grcuda_mm_cc* h, long long* clamped, long long* overflow) {
  GRB_CUDA(cudaDeviceSynchronize());
  std::vector<MMCCChan> st((size_t)h->nchan);
  GRB_CUDA(cudaMemcpy(st.data(), h->d_chan.p, st.size() * sizeof(MMCCChan), cudaMemcpyDeviceToHost));
  long long c = 0, o = 0;
  for (auto& s : st) { c += s.clamped; o += s.overflow; }
  if (clamped) *clamped = c;
  if (overflow) *overflow = o;
  return GRCUDA_OK;
}
int grcuda_clock_recovery_mm_cc_work_device(grcuda_mm_cc* h, long ninput_rows, long abs_row0, const grcuda_complex* d_in,
                                            grcuda_complex* d_out, float* d_err, int max_out, int* d_counts, void* stream) {
  if (max_out < 0) return set_error(GRCUDA_EINVAL, "clock_recovery_mm_cc: max_out < 0");
  return h->launch((const float2*)d_in, ninput_rows, abs_row0, (float2*)d_out, d_err, max_out, d_counts ? d_counts : h->d_counts.as<int>(),
                   h->pick(stream));
}
int grcuda_clock_recovery_mm_cc_work(grcuda_mm_cc* h, int noutput_items, int ninput_items, const grcuda_complex* in, grcuda_complex* out,
                                     float* err_out, int* consumed) {
  if (consumed) *consumed = 0;
  if (h->nchan != 1) return set_error(GRCUDA_EINVAL, "host work() is the single-stream form (nchan == 1)");
  if (noutput_items <= 0 || ninput_items <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)ninput_items * sizeof(float2))) || (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float2))) ||
      (err_out && (rc = h->d_err.reserve((size_t)noutput_items * sizeof(float)))))
    return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)ninput_items * sizeof(float2), h->stream))) return rc;
  // the runtime re-presents unconsumed items at in[0]: the loop's position is 0 relative to this call
  const long long zero = 0;
  GRB_CUDA(cudaMemcpyAsync((char*)h->d_chan.p + offsetof(MMCCChan, next_abs), &zero, sizeof zero, cudaMemcpyHostToDevice, h->stream));
  if ((rc = h->launch(h->d_in.as<float2>(), ninput_items, 0, h->d_out.as<float2>(), err_out ? h->d_err.as<float>() : nullptr, noutput_items,
                      h->d_counts.as<int>(), h->stream)))
    return rc;
  int produced = 0;
  MMCCChan st;
  GRB_CUDA(cudaMemcpyAsync(&produced, h->d_counts.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GRB_CUDA(cudaMemcpyAsync(&st, h->d_chan.p, sizeof st, cudaMemcpyDeviceToHost, h->stream));
  GRB_CUDA(cudaStreamSynchronize(h->stream));
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)produced * sizeof(float2), h->stream))) return rc;
  if (err_out && (rc = h->stager.d2h(err_out, h->d_err.p, (size_t)produced * sizeof(float), h->stream))) return rc;
  if (consumed) *consumed = (int)st.next_abs;   // :207-214: consume_each(ii)
  return produced;
}

}  // extern "C"
