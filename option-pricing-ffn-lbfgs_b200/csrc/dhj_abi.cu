// dhj_abi.cu — host side of libdhj.so: contexts, device/pinned buffers, option books, and the extern "C"
// entry points declared in include/dhj.h.  No torch, no Python: plain CUDA runtime.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <omp.h>

#include "dhj.h"
#include "dhj_kernels.cuh"
#include "dhj_generate.cuh"

using namespace dhj;

extern "C" int dhj_host_threads();      // dhj_lbfgs.cpp: threads of the host-side parallel loops

namespace {

char g_init_error[512] = "";

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max(bytes, (size_t)256);
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = std::max(bytes, (size_t)256);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Host description of the option slices (options grouped by maturity, first-appearance order).
struct BookHost {
  int M = 0, n_slices = 0, max_slice = 0;
  std::vector<double> slice_T;
  std::vector<int> slice_off, pos;
  std::vector<unsigned char> call;
  void group(const double* maturity, const int32_t* is_call, int m) {
    M = m;
    slice_T.clear();
    std::vector<int> slice_of(m);
    for (int o = 0; o < m; ++o) {
      int s = -1;
      for (size_t t = 0; t < slice_T.size(); ++t)
        if (slice_T[t] == maturity[o] || (std::isnan(slice_T[t]) && std::isnan(maturity[o]))) { s = (int)t; break; }
      if (s < 0) { s = (int)slice_T.size(); slice_T.push_back(maturity[o]); }
      slice_of[o] = s;
    }
    n_slices = (int)slice_T.size();
    slice_off.assign(n_slices + 1, 0);
    for (int o = 0; o < m; ++o) slice_off[slice_of[o] + 1]++;
    max_slice = 0;
    for (int s = 0; s < n_slices; ++s) { max_slice = std::max(max_slice, slice_off[s + 1]); slice_off[s + 1] += slice_off[s]; }
    pos.assign(m, 0); call.assign(m, 0);
    std::vector<int> fill(slice_off.begin(), slice_off.end() - 1);
    for (int o = 0; o < m; ++o) {
      int d = fill[slice_of[o]]++;
      pos[d] = o;
      call[d] = is_call[o] ? 1 : 0;
    }
  }
  void grid(const double* maturities, int nT, int nK, int is_call_flag) {
    M = nT * nK; n_slices = nT; max_slice = nK;
    slice_T.assign(maturities, maturities + nT);
    slice_off.resize(nT + 1);
    for (int t = 0; t <= nT; ++t) slice_off[t] = t * nK;
    pos.resize(M); call.assign(M, is_call_flag ? 1 : 0);
    for (int o = 0; o < M; ++o) pos[o] = o;
  }
  // bytes of the packed device image: slice_T | slice_off | pos | call (8-byte aligned segments)
  size_t packed_bytes() const {
    return align8(n_slices * sizeof(double)) + align8((n_slices + 1) * sizeof(int)) + align8(M * sizeof(int)) +
           align8(M);
  }
  static size_t align8(size_t b) { return (b + 7) & ~(size_t)7; }
  void pack(unsigned char* dst) const {
    size_t o = 0;
    memcpy(dst + o, slice_T.data(), n_slices * sizeof(double)); o += align8(n_slices * sizeof(double));
    memcpy(dst + o, slice_off.data(), (n_slices + 1) * sizeof(int)); o += align8((n_slices + 1) * sizeof(int));
    memcpy(dst + o, pos.data(), M * sizeof(int)); o += align8(M * sizeof(int));
    memcpy(dst + o, call.data(), M);
  }
  void view(SliceView* v, const unsigned char* dbase) const {
    size_t o = 0;
    v->n_slices = n_slices; v->n_options = M;
    v->slice_T = (const double*)(dbase + o); o += align8(n_slices * sizeof(double));
    v->slice_off = (const int*)(dbase + o); o += align8((n_slices + 1) * sizeof(int));
    v->pos = (const int*)(dbase + o); o += align8(M * sizeof(int));
    v->call = dbase + o;
  }
};

// Option books (packed slice tables + shared strike row) live in a small ring of device slots, so that a call
// with a new (K, T) table neither waits for the device nor disturbs launches that still read an older table
// (DoubleHeston(...).pricing() changes the book on every call; the previous design synchronised the whole device).
constexpr int kBookSlots = 8;
struct BookSlot {
  DevBuf d;
  PinBuf h;
  std::vector<unsigned char> image;     // what the slot holds (empty: nothing)
  cudaEvent_t uploaded = nullptr;        // the H2D copy of `image`
  cudaEvent_t last_use = nullptr;        // the last launch that reads the slot
  cudaStream_t upload_stream = nullptr;
};

// Chunks of a host-buffer call rotate through three slots (stream + device buffers + pinned staging each): a slot's
// upload, kernel and download are serialised by its stream, so with two slots a chunk every (H2D + kernel + D2H) / 2
// is the best the pipeline can do — slower than the kernel as soon as the host link gives a rank less than ~20 GB/s
// (eight ranks copying at once on one host: 11.5 GB/s down per GPU, measured; profiles/README.md r02).
constexpr int kSlots = 3;
struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  DevBuf d_in, d_out;
  PinBuf h_in, h_out;
  // deferred copy-out of a staged chunk
  double* user_out = nullptr;
  size_t out_bytes = 0;
  bool pending = false;
  // generator path: several destination arrays per chunk
  struct Piece { void* dst; size_t off, bytes; };
  std::vector<Piece> pieces;
  void drain_pieces(void (*copy)(void*, const void*, size_t)) {
    for (const Piece& pc : pieces) copy(pc.dst, (const unsigned char*)h_out.p + pc.off, pc.bytes);
    pieces.clear();
  }
};

}  // namespace

struct dhj_ctx {
  int device = 0;
  int sm_count = 0;
  int loss_blocks_per_sm = 1;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  char err[512] = "";
  // cached option books for the pricing entry points
  BookSlot books[kBookSlots];
  int book_next = 0;                          // ring position of the next slot to overwrite
  int book_cur = 0;                           // slot of the last upload_book call
  Slot slots[kSlots];
  // loss path
  DevBuf d_x, d_xv, d_idx, d_f, d_fg, d_counters, d_prices;
  PinBuf h_x, h_res;
  size_t counters_zeroed = 0;
  DevBuf d_peak;
  // device buffers of destroyed markets, kept for the next dhj_market_create: cudaMalloc / cudaFree synchronise the
  // whole device, which stalls the other lock-step pipelines of a calibrate_many call every time a market comes or goes
  std::vector<DevBuf> pool;
  cudaError_t pool_take(DevBuf* out, size_t bytes) {
    int best = -1;
    for (size_t i = 0; i < pool.size(); ++i)
      if (pool[i].cap >= bytes && (best < 0 || pool[i].cap < pool[(size_t)best].cap)) best = (int)i;
    if (best >= 0 && pool[(size_t)best].cap <= 4 * std::max(bytes, (size_t)4096)) {
      *out = pool[(size_t)best];
      pool.erase(pool.begin() + best);
      return cudaSuccess;
    }
    *out = DevBuf();
    return out->reserve(bytes);
  }
  void pool_give(DevBuf* b) {
    if (b->p && pool.size() < 32 && b->cap <= ((size_t)256 << 20)) pool.push_back(*b);
    else b->release();
    b->p = nullptr; b->cap = 0;
  }
};

struct dhj_market {
  dhj_ctx* ctx = nullptr;
  int n_markets = 0, M = 0, N = 128;
  double r = 0.0;
  BookHost book;
  DevBuf d_book, d_strike, d_S0, d_price;
  long long strike_stride = 0;
  SliceView view{};
};

namespace {

int fail(dhj_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_init_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define DHJ_CUDA(ctx, call)                                                                          \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail((ctx), e__ == cudaErrorMemoryAllocation ? DHJ_ERR_NOMEM : DHJ_ERR_CUDA,            \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);     \
  } while (0)

// staging copies between the caller's pageable memory and the pinned slots: a single thread moves ~10 GB/s, less
// than the kernel consumes per chunk, so large copies are split over eight host threads (four: 14.75 ms per
// 1 Mi-set grid from pageable memory, eight: 14.28)
int copy_threads() { return std::max(1, std::min(8, dhj_host_threads())); }

void host_copy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kPiece = (size_t)1 << 18;
  if (bytes < 4 * kPiece) { memcpy(dst, src, bytes); return; }
  const long long pieces = (long long)((bytes + kPiece - 1) / kPiece);
#pragma omp parallel for schedule(static) num_threads(copy_threads())
  for (long long i = 0; i < pieces; ++i) {
    const size_t off = (size_t)i * kPiece;
    memcpy((char*)dst + off, (const char*)src + off, std::min(kPiece, bytes - off));
  }
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// Batch size of the batch engine for a launch of `n_units` units of `items_per_unit` items each (pricing: unit = item).
// A block's time is its phase 1 (one warp, ~0.9 item times) plus that of its busiest warp (items are dealt to the four
// warps in turn); a launch's time is the number of waves of resident blocks times that.  Minimise
// waves x (0.9 + ceil(items per batch / 4)) over whole units per batch; ties go to the larger batch.  E.g. 7 000 loss
// evaluations of a 3-slice market: 10 units per batch = 700 blocks = 2 waves (the second nearly empty) of 8 items per
// warp, 4 units per batch = 3 waves of 3.  Large launches end at full batches, a single calibration at one unit per
// block (every unit on its own SM: latency).
int pick_units_per_batch(long long n_units, int items_per_unit, int max_items, long long resident_blocks) {
  const int upb_max = std::max(1, max_items / std::max(1, items_per_unit));
  long long best_cost = -1;
  int best = 1;
  for (int upb = upb_max; upb >= 1; --upb) {
    const long long blocks = (n_units + upb - 1) / upb;
    const long long waves = (blocks + resident_blocks - 1) / resident_blocks;
    const long long cost = waves * (9 + 10 * (((long long)upb * items_per_unit + kBatchWarps - 1) / kBatchWarps));
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = upb; }
  }
  return best;
}

// k_price_batch handles slices of <= 8 strikes (warp per item, lane per k, <= 32 items per block); k_price_dense the rest
int launch_price(dhj_ctx* ctx, const SliceView& v, PriceArgs a, int max_slice, cudaStream_t st);

int launch_price(dhj_ctx* ctx, const SliceView& v, PriceArgs a, int max_slice, cudaStream_t st) {
  const long long items = a.P * (long long)v.n_slices;
  if (max_slice <= kBatchMaxStrikes) {
    // full batches of 32 items unless the launch is only a few waves long (mid-size loss rounds through the split path)
    const long long resident = (long long)ctx->sm_count * ctx->loss_blocks_per_sm;
    a.items_per_batch = (items > 64 * resident * kPriceItems) ? kPriceItems
                                                              : pick_units_per_batch(items, 1, kPriceItems, resident);
    const long long batches = (items + a.items_per_batch - 1) / a.items_per_batch;
    // one block per batch: the hardware block scheduler balances the SMs dynamically (a persistent grid with a
    // static batch -> block map measured ~3 % slower: the slowest SM sets the time)
    const long long cap = 2147483647LL;
    // DHJ_DEBUG_EXTRA_SMEM (bytes): developer knob that pads the launch with unused dynamic shared memory to
    // lower the resident block count (occupancy experiments, profiles/README.md); never set in production
    static const size_t extra = [] {
      const char* e = getenv("DHJ_DEBUG_EXTRA_SMEM");
      size_t b = e ? (size_t)atol(e) : 0;
      if (b) cudaFuncSetAttribute(k_price_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
      return b;
    }();
    const int grid = (int)std::max<long long>(1, std::min(batches, cap));
    if (a.items_per_batch == kPriceItems) k_price_batch<true><<<grid, kBatchThreads, extra, st>>>(v, a);
    else k_price_batch<false><<<grid, kBatchThreads, 0, st>>>(v, a);
  } else {
    // a warp per item, four items per block; the block's shared memory (four private strike tables) exceeds the
    // 48 KB static limit: opt in once
    static const cudaError_t attr = cudaFuncSetAttribute(k_price_dense, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)sizeof(DenseSmem));
    DHJ_CUDA(ctx, attr);
    const long long blocks = (items + kDenseWarps - 1) / kDenseWarps;
    k_price_dense<<<(int)std::max<long long>(1, std::min(blocks, 2147483647LL)), 32 * kDenseWarps, sizeof(DenseSmem),
                    st>>>(v, a);
  }
  DHJ_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return DHJ_OK;
}

// Find (or upload into the ring) the packed book image + shared strike table for a pricing call on `stream`.
// No device-wide synchronisation: a new table goes to the least recently filled slot (after the last launch that
// read that slot has finished — normally long ago), its copy is ordered before the caller's launches by stream
// order or an event wait, and launches that still read other slots are not disturbed.
int upload_book(dhj_ctx* ctx, const BookHost& book, const double* shared_strikes, int n_shared,
                cudaStream_t stream, SliceView* v) {
  const size_t bb = book.packed_bytes();
  const size_t sb = BookHost::align8((size_t)n_shared * sizeof(double));
  std::vector<unsigned char> image(bb + sb);
  book.pack(image.data());
  if (n_shared) memcpy(image.data() + bb, shared_strikes, (size_t)n_shared * sizeof(double));
  int hit = -1;
  for (int i = 0; i < kBookSlots && hit < 0; ++i)
    if (ctx->books[i].image == image) hit = i;
  if (hit < 0) {
    hit = ctx->book_next;
    ctx->book_next = (ctx->book_next + 1) % kBookSlots;
    BookSlot& b = ctx->books[hit];
    b.image.clear();
    // the slot's device table may still be read, its pinned staging may still feed the previous upload
    DHJ_CUDA(ctx, cudaEventSynchronize(b.last_use));
    DHJ_CUDA(ctx, cudaEventSynchronize(b.uploaded));
    DHJ_CUDA(ctx, b.d.reserve(image.size()));
    DHJ_CUDA(ctx, b.h.reserve(image.size()));
    memcpy(b.h.p, image.data(), image.size());
    DHJ_CUDA(ctx, cudaMemcpyAsync(b.d.p, b.h.p, image.size(), cudaMemcpyHostToDevice, stream));
    DHJ_CUDA(ctx, cudaEventRecord(b.uploaded, stream));
    b.upload_stream = stream;
    b.image.swap(image);
  } else if (ctx->books[hit].upload_stream != stream) {
    DHJ_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->books[hit].uploaded, 0));
  }
  ctx->book_cur = hit;
  const BookSlot& b = ctx->books[hit];
  book.view(v, (const unsigned char*)b.d.p);
  v->strike = n_shared ? (const double*)((const unsigned char*)b.d.p + bb) : nullptr;
  v->strike_stride = 0;
  return DHJ_OK;
}

// the launches of the current call on `stream` read the current book slot
int book_used(dhj_ctx* ctx, cudaStream_t stream) {
  DHJ_CUDA(ctx, cudaEventRecord(ctx->books[ctx->book_cur].last_use, stream));
  return DHJ_OK;
}

int check_common(dhj_ctx* ctx, const void* params, int64_t P, const void* S0, int32_t N, double L, const void* out) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!params || !S0 || !out) return fail(ctx, DHJ_ERR_ARG, "null array argument");
  if (P < 0) return fail(ctx, DHJ_ERR_ARG, "P must be >= 0 (got %lld)", (long long)P);
  if (N < 1) return fail(ctx, DHJ_ERR_ARG, "N must be >= 1 (got %d)", N);
  if (!(L == L)) return fail(ctx, DHJ_ERR_ARG, "L is NaN");
  return DHJ_OK;
}

// Chunked, double-buffered host-to-host pricing: params/S0/strike rows in, price rows out.
int run_price_host(dhj_ctx* ctx, SliceView v, int max_slice, const double* params, int64_t P, const double* S0,
                   int64_t s0_stride, const double* strike, int64_t strike_stride, double* out) {
  const int M = v.n_options;
  if (P == 0 || M == 0) return DHJ_OK;
  const size_t in_row = (size_t)kNumParams + (s0_stride ? 1 : 0) + (strike_stride ? (size_t)M : 0);
  const size_t row_doubles = in_row + (size_t)M;
  int64_t chunk = (int64_t)std::max<size_t>(256, std::min<size_t>(131072, ((size_t)1 << 23) / row_doubles));
  // a call that moves more than 16 MB is cut into at least 8 chunks so that transfers overlap kernels even when
  // one set is large (the 200 x 20 surface: 32 KB of prices per set, 1 024 sets -> 8 chunks of 128)
  if ((size_t)P * row_doubles * sizeof(double) > ((size_t)16 << 20))
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(32, (P + 7) / 8));
  const bool pin_in = is_pinned_host(params) && (!s0_stride || is_pinned_host(S0)) &&
                      (!strike_stride || is_pinned_host(strike));
  const bool pin_out = is_pinned_host(out);
  // Chunk schedule: the first chunk's upload and the last chunk's download cannot overlap any kernel, so with
  // pinned buffers the schedule ramps up from chunk/8 and down to chunk/8 again (exposed transfer 0.54 -> 0.07 ms
  // on the 1 Mi-set grid: 13.87 -> 13.23 ms).  Pageable buffers keep uniform chunks: their staging copies make
  // the host the bottleneck while the chunks are small (14.3 vs 14.9 ms).
  std::vector<int64_t> sizes;
  {
    std::vector<int64_t> head, tail;
    int64_t left = P;
    if (pin_in && pin_out && chunk >= 1024)
      for (int64_t s = chunk / 8; s < chunk && left > 2 * s + chunk; s *= 2) {
        head.push_back(s); tail.push_back(s);
        left -= 2 * s;
      }
    sizes = head;
    for (; left > 0; left -= chunk) sizes.push_back(std::min(chunk, left));
    sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
  }
  // A previous call on this context may have failed half-way (DHJ_ERR_NOMEM is recoverable) and left a staged
  // chunk behind: its caller's buffer is gone, so nothing of it may be copied out now.  Drain and forget.
  for (int i = 0; i < kSlots; ++i) {
    Slot& sl = ctx->slots[i];
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    sl.pending = false; sl.user_out = nullptr; sl.out_bytes = 0; sl.pieces.clear();
    // the option book was uploaded on another stream
    DHJ_CUDA(ctx, cudaStreamWaitEvent(sl.stream, ctx->books[ctx->book_cur].uploaded, 0));
  }
  // DHJ_DEBUG_FAIL_AT_CHUNK (test hook, tests/test_gpu_dropin.py): return DHJ_ERR_NOMEM before that chunk is staged,
  // as a failed allocation would, with earlier chunks still in flight
  const char* fail_env = getenv("DHJ_DEBUG_FAIL_AT_CHUNK");
  const long fail_at = fail_env ? atol(fail_env) : -1;
  int slot_i = 0;
  int64_t lo = 0;
  for (size_t ci = 0; ci < sizes.size(); lo += sizes[ci], ++ci, slot_i = (slot_i + 1) % kSlots) {
    const int64_t n = sizes[ci];
    Slot& sl = ctx->slots[slot_i];
    if ((long)ci == fail_at) return fail(ctx, DHJ_ERR_NOMEM, "injected failure at chunk %ld (DHJ_DEBUG_FAIL_AT_CHUNK)", fail_at);
    // the slot's previous chunk must have left its buffers
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    if (sl.pending) { host_copy(sl.user_out, sl.h_out.p, sl.out_bytes); sl.pending = false; }
    const size_t pb = (size_t)n * kNumParams * sizeof(double);
    const size_t s0b = s0_stride ? (size_t)n * sizeof(double) : sizeof(double);
    const size_t kb = strike_stride ? (size_t)n * M * sizeof(double) : 0;
    const size_t ob = (size_t)n * M * sizeof(double);
    DHJ_CUDA(ctx, sl.d_in.reserve(pb + s0b + kb));
    DHJ_CUDA(ctx, sl.d_out.reserve(ob));
    unsigned char* din = (unsigned char*)sl.d_in.p;
    const double* src_params = params + lo * kNumParams;
    const double* src_s0 = s0_stride ? S0 + lo : S0;
    const double* src_strike = strike_stride ? strike + lo * (int64_t)M : nullptr;
    if (pin_in) {
      DHJ_CUDA(ctx, cudaMemcpyAsync(din, src_params, pb, cudaMemcpyHostToDevice, sl.stream));
      DHJ_CUDA(ctx, cudaMemcpyAsync(din + pb, src_s0, s0b, cudaMemcpyHostToDevice, sl.stream));
      if (kb) DHJ_CUDA(ctx, cudaMemcpyAsync(din + pb + s0b, src_strike, kb, cudaMemcpyHostToDevice, sl.stream));
    } else {
      DHJ_CUDA(ctx, sl.h_in.reserve(pb + s0b + kb));
      unsigned char* hin = (unsigned char*)sl.h_in.p;
      host_copy(hin, src_params, pb);
      host_copy(hin + pb, src_s0, s0b);
      if (kb) host_copy(hin + pb + s0b, src_strike, kb);
      DHJ_CUDA(ctx, cudaMemcpyAsync(din, hin, pb + s0b + kb, cudaMemcpyHostToDevice, sl.stream));
    }
    SliceView vv = v;
    if (strike_stride) { vv.strike = (const double*)(din + pb + s0b); vv.strike_stride = M; }
    PriceArgs a;
    a.params = (const double*)din; a.S0 = (const double*)(din + pb); a.s0_stride = s0_stride ? 1 : 0;
    a.row_index = nullptr; a.P = n; a.transform = 0; a.out = (double*)sl.d_out.p;
    int rc = launch_price(ctx, vv, a, max_slice, sl.stream);
    if (rc) return rc;
    double* dst = out + lo * (int64_t)M;
    if (pin_out) {
      DHJ_CUDA(ctx, cudaMemcpyAsync(dst, sl.d_out.p, ob, cudaMemcpyDeviceToHost, sl.stream));
    } else {
      DHJ_CUDA(ctx, sl.h_out.reserve(ob));
      DHJ_CUDA(ctx, cudaMemcpyAsync(sl.h_out.p, sl.d_out.p, ob, cudaMemcpyDeviceToHost, sl.stream));
      sl.user_out = dst; sl.out_bytes = ob; sl.pending = true;
    }
    DHJ_CUDA(ctx, cudaEventRecord(sl.done, sl.stream));
  }
  for (int i = 0; i < kSlots; ++i) {
    Slot& sl = ctx->slots[i];
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    if (sl.pending) { host_copy(sl.user_out, sl.h_out.p, sl.out_bytes); sl.pending = false; }
  }
  return DHJ_OK;
}

int stage_x(dhj_ctx* ctx, const dhj_market* mk, const double* x, const int32_t* market_index, int64_t B,
            const int** d_index) {
  if (!ctx || !mk) return fail(ctx, DHJ_ERR_ARG, "null context or market");
  if (mk->ctx != ctx) return fail(ctx, DHJ_ERR_ARG, "market belongs to another context");
  if (!x) return fail(ctx, DHJ_ERR_ARG, "null x");
  if (B < 0 || B > (int64_t)100000000) return fail(ctx, DHJ_ERR_ARG, "bad batch size %lld", (long long)B);
  const size_t xb = (size_t)B * kNumParams * sizeof(double);
  const size_t ib = market_index ? (size_t)B * sizeof(int) : 0;
  if (market_index)
    for (int64_t i = 0; i < B; ++i)
      if (market_index[i] < 0 || market_index[i] >= mk->n_markets)
        return fail(ctx, DHJ_ERR_ARG, "market_index[%lld] = %d out of range [0,%d)", (long long)i,
                    market_index[i], mk->n_markets);
  DHJ_CUDA(ctx, ctx->h_x.reserve(xb + ib));
  DHJ_CUDA(ctx, ctx->d_x.reserve(xb + ib));
  host_copy(ctx->h_x.p, x, xb);
  if (ib) memcpy((unsigned char*)ctx->h_x.p + xb, market_index, ib);
  DHJ_CUDA(ctx, cudaMemcpyAsync(ctx->d_x.p, ctx->h_x.p, xb + ib, cudaMemcpyHostToDevice, ctx->stream));
  *d_index = ib ? (const int*)((unsigned char*)ctx->d_x.p + xb) : nullptr;
  return DHJ_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int dhj_abi_version(void) { return DHJ_ABI_VERSION; }

int dhj_device_count(int* count) {
  if (!count) return fail(nullptr, DHJ_ERR_ARG, "null count");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return fail(nullptr, DHJ_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  return DHJ_OK;
}

int dhj_init(int device, dhj_ctx** out) {
  if (!out) return fail(nullptr, DHJ_ERR_ARG, "null ctx pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, DHJ_ERR_CUDA, "no CUDA device available (%s); libdhj has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) return fail(nullptr, DHJ_ERR_ARG, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  DHJ_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, DHJ_ERR_CUDA, "device %d is sm_%d%d; libdhj is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  DHJ_CUDA(nullptr, cudaSetDevice(device));
  dhj_ctx* ctx = new (std::nothrow) dhj_ctx();
  if (!ctx) return fail(nullptr, DHJ_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  for (int i = 0; i < kBookSlots && e2 == cudaSuccess; ++i) {
    e2 = cudaEventCreateWithFlags(&ctx->books[i].uploaded, cudaEventDisableTiming);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->books[i].last_use, cudaEventDisableTiming);
  }
  for (int i = 0; i < kSlots && e2 == cudaSuccess; ++i) {
    e2 = cudaStreamCreateWithFlags(&ctx->slots[i].stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->slots[i].done, cudaEventDisableTiming);
    if (e2 == cudaSuccess) e2 = cudaEventRecord(ctx->slots[i].done, ctx->slots[i].stream);
  }
  int bps = 0;
  if (e2 == cudaSuccess)
    e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_loss_batch, kBatchThreads, 0);
  if (e2 != cudaSuccess) {
    fail(nullptr, DHJ_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e2));
    dhj_destroy(ctx);
    return DHJ_ERR_CUDA;
  }
  ctx->loss_blocks_per_sm = std::max(1, bps);
  *out = ctx;
  return DHJ_OK;
}

int dhj_destroy(dhj_ctx* ctx) {
  if (!ctx) return DHJ_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = ctx->slots[i];
    s.d_in.release(); s.d_out.release(); s.h_in.release(); s.h_out.release();
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (int i = 0; i < kBookSlots; ++i) {
    BookSlot& b = ctx->books[i];
    b.d.release(); b.h.release();
    if (b.uploaded) cudaEventDestroy(b.uploaded);
    if (b.last_use) cudaEventDestroy(b.last_use);
  }
  ctx->d_x.release(); ctx->d_xv.release(); ctx->d_idx.release(); ctx->d_f.release(); ctx->d_fg.release();
  ctx->d_counters.release(); ctx->d_prices.release(); ctx->h_x.release(); ctx->h_res.release();
  ctx->d_peak.release();
  for (DevBuf& b : ctx->pool) b.release();
  ctx->pool.clear();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return DHJ_OK;
}

const char* dhj_last_error(const dhj_ctx* ctx) { return ctx ? ctx->err : g_init_error; }

int dhj_launch_count(const dhj_ctx* ctx, int64_t* count) {
  if (!ctx || !count) return DHJ_ERR_ARG;
  *count = ctx->launches;
  return DHJ_OK;
}

// ---- pricing ---------------------------------------------------------------------------------
int dhj_price_list(dhj_ctx* ctx, const double* params, int64_t P, const double* S0, int64_t s0_stride,
                   double r, double q, const double* strike, int64_t strike_stride, const double* maturity,
                   const int32_t* is_call, int32_t M, int32_t N, double L, double* out) {
  int rc = check_common(ctx, params, P, S0, N, L, out);
  if (rc) return rc;
  if (M < 0) return fail(ctx, DHJ_ERR_ARG, "M must be >= 0");
  if (M == 0 || P == 0) return DHJ_OK;
  if (!strike || !maturity || !is_call) return fail(ctx, DHJ_ERR_ARG, "null option table");
  if (s0_stride != 0 && s0_stride != 1) return fail(ctx, DHJ_ERR_ARG, "s0_stride must be 0 or 1");
  if (strike_stride != 0 && strike_stride != M) return fail(ctx, DHJ_ERR_ARG, "strike_stride must be 0 or M");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  BookHost book;
  book.group(maturity, is_call, M);
  SliceView v;
  rc = upload_book(ctx, book, strike_stride ? nullptr : strike, strike_stride ? 0 : M, ctx->stream, &v);
  if (rc) return rc;
  v.scale_by_spot = 0; v.n_cos = N; v.r = r; v.q = q; v.L = L;
  return run_price_host(ctx, v, book.max_slice, params, P, S0, s0_stride, strike, strike_stride, out);
}

static int grid_view(dhj_ctx* ctx, const double* strikes, int32_t nK, const double* maturities, int32_t nT,
                     int32_t scale_by_spot, int32_t is_call, int32_t N, double r, double q, double L,
                     cudaStream_t stream, SliceView* v) {
  if (nK < 1 || nT < 1) return fail(ctx, DHJ_ERR_ARG, "nK and nT must be >= 1");
  if (!strikes || !maturities) return fail(ctx, DHJ_ERR_ARG, "null grid table");
  BookHost book;
  book.grid(maturities, nT, nK, is_call);
  std::vector<double> tiled((size_t)nT * nK);
  for (int t = 0; t < nT; ++t) memcpy(tiled.data() + (size_t)t * nK, strikes, (size_t)nK * sizeof(double));
  int rc = upload_book(ctx, book, tiled.data(), nT * nK, stream, v);
  if (rc) return rc;
  v->scale_by_spot = scale_by_spot ? 1 : 0; v->n_cos = N; v->r = r; v->q = q; v->L = L;
  return DHJ_OK;
}

int dhj_price_grid(dhj_ctx* ctx, const double* params, int64_t P, const double* S0, int64_t s0_stride,
                   double r, double q, const double* strikes, int32_t nK, const double* maturities,
                   int32_t nT, int32_t scale_by_spot, int32_t is_call, int32_t N, double L, double* out) {
  int rc = check_common(ctx, params, P, S0, N, L, out);
  if (rc) return rc;
  if (s0_stride != 0 && s0_stride != 1) return fail(ctx, DHJ_ERR_ARG, "s0_stride must be 0 or 1");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  SliceView v;
  rc = grid_view(ctx, strikes, nK, maturities, nT, scale_by_spot, is_call, N, r, q, L, ctx->stream, &v);
  if (rc) return rc;
  if (P == 0) return DHJ_OK;
  return run_price_host(ctx, v, nK, params, P, S0, s0_stride, nullptr, 0, out);
}

int dhj_price_grid_dev(dhj_ctx* ctx, const double* d_params, int64_t P, const double* d_S0,
                       int64_t s0_stride, double r, double q, const double* strikes, int32_t nK,
                       const double* maturities, int32_t nT, int32_t scale_by_spot, int32_t is_call,
                       int32_t N, double L, double* d_out, void* stream) {
  int rc = check_common(ctx, d_params, P, d_S0, N, L, d_out);
  if (rc) return rc;
  if (s0_stride != 0 && s0_stride != 1) return fail(ctx, DHJ_ERR_ARG, "s0_stride must be 0 or 1");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;   // NULL = the default stream
  SliceView v;
  rc = grid_view(ctx, strikes, nK, maturities, nT, scale_by_spot, is_call, N, r, q, L, st, &v);
  if (rc) return rc;
  if (P == 0) return DHJ_OK;
  PriceArgs a;
  a.params = d_params; a.S0 = d_S0; a.s0_stride = s0_stride; a.row_index = nullptr; a.P = P; a.transform = 0;
  a.out = d_out;
  rc = launch_price(ctx, v, a, nK, st);
  if (rc) return rc;
  return book_used(ctx, st);
}

// ---- calibration loss ------------------------------------------------------------------------
int dhj_market_create(dhj_ctx* ctx, int32_t n_markets, int32_t M, const double* S0, double r,
                      const double* strike, int64_t strike_stride, const double* maturity,
                      const int32_t* is_call, const double* price, int32_t N, dhj_market** out) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!out) return fail(ctx, DHJ_ERR_ARG, "null market pointer");
  *out = nullptr;
  if (n_markets < 1 || M < 1) return fail(ctx, DHJ_ERR_ARG, "n_markets and M must be >= 1");
  if (!S0 || !strike || !maturity || !is_call || !price) return fail(ctx, DHJ_ERR_ARG, "null market array");
  if (N < 1) return fail(ctx, DHJ_ERR_ARG, "N must be >= 1");
  if (strike_stride != 0 && strike_stride != M) return fail(ctx, DHJ_ERR_ARG, "strike_stride must be 0 or M");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  dhj_market* mk = new (std::nothrow) dhj_market();
  if (!mk) return fail(ctx, DHJ_ERR_NOMEM, "out of host memory");
  mk->ctx = ctx; mk->n_markets = n_markets; mk->M = M; mk->N = N; mk->r = r;
  mk->strike_stride = strike_stride;
  mk->book.group(maturity, is_call, M);
  const size_t bb = mk->book.packed_bytes();
  std::vector<unsigned char> image(bb);
  mk->book.pack(image.data());
  const size_t kbytes = (size_t)(strike_stride ? n_markets : 1) * M * sizeof(double);
  const size_t pbytes = (size_t)n_markets * M * sizeof(double);
  cudaError_t e = ctx->pool_take(&mk->d_book, bb);
  if (e == cudaSuccess) e = ctx->pool_take(&mk->d_strike, kbytes);
  if (e == cudaSuccess) e = ctx->pool_take(&mk->d_S0, (size_t)n_markets * sizeof(double));
  if (e == cudaSuccess) e = ctx->pool_take(&mk->d_price, pbytes);
  if (e == cudaSuccess) e = cudaMemcpy(mk->d_book.p, image.data(), bb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mk->d_strike.p, strike, kbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mk->d_S0.p, S0, (size_t)n_markets * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mk->d_price.p, price, pbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    dhj_market_destroy(mk);
    return fail(ctx, e == cudaErrorMemoryAllocation ? DHJ_ERR_NOMEM : DHJ_ERR_CUDA, "market upload failed: %s",
                cudaGetErrorString(e));
  }
  mk->book.view(&mk->view, (const unsigned char*)mk->d_book.p);
  mk->view.strike = (const double*)mk->d_strike.p;
  mk->view.strike_stride = strike_stride;
  mk->view.scale_by_spot = 0; mk->view.n_cos = N; mk->view.r = r; mk->view.q = 0.0; mk->view.L = 10.0;
  *out = mk;
  return DHJ_OK;
}

int dhj_market_destroy(dhj_market* mk) {
  if (!mk) return DHJ_OK;
  if (mk->ctx) {
    cudaSetDevice(mk->ctx->device);
    // launches that read the market's tables are all on the context's stream: once it has drained the buffers can
    // serve the next market
    cudaStreamSynchronize(mk->ctx->stream);
    mk->ctx->pool_give(&mk->d_book); mk->ctx->pool_give(&mk->d_strike);
    mk->ctx->pool_give(&mk->d_S0); mk->ctx->pool_give(&mk->d_price);
  } else {
    mk->d_book.release(); mk->d_strike.release(); mk->d_S0.release(); mk->d_price.release();
  }
  delete mk;
  return DHJ_OK;
}

static int run_loss(dhj_ctx* ctx, const dhj_market* mk, const double* x, const int32_t* market_index, int64_t B,
                    int fd, double h, double* out_f, double* out_g, double* out_f_all) {
  const int* d_index = nullptr;
  DHJ_CUDA(ctx, cudaSetDevice(ctx ? ctx->device : 0));
  int rc = stage_x(ctx, mk, x, market_index, B, &d_index);
  if (rc) return rc;
  if (!out_f || (fd && !out_g)) return fail(ctx, DHJ_ERR_ARG, "null output");
  if (B == 0) return DHJ_OK;
  const int per = fd ? kFdPoints : 1;
  const int64_t n_units = B * per;
  if (n_units > 2000000000LL) return fail(ctx, DHJ_ERR_ARG, "batch too large for one launch");
  DHJ_CUDA(ctx, ctx->d_f.reserve((size_t)n_units * sizeof(double)));
  size_t res_bytes = (size_t)B * sizeof(double);
  if (fd) {
    res_bytes = (size_t)n_units * sizeof(double);
    DHJ_CUDA(ctx, ctx->d_fg.reserve(res_bytes));
  }
  const SliceView& v = mk->view;
  // Large batches go through the pricing kernel (expand -> k_price_batch -> reduce): its batches of 32 items keep the
  // four warps balanced whatever the number of slices per loss evaluation, which pays for two extra small launches
  // from a few thousand evaluations on.  Both paths run the same arithmetic and produce the same bits
  // (DHJ_DEBUG_SPLIT_UNITS: developer knob for measuring the crossover).
  static const int64_t kSplitUnits = [] {
    const char* e = getenv("DHJ_DEBUG_SPLIT_UNITS");
    return e ? (int64_t)atoll(e) : (int64_t)8192;
  }();
  if (mk->book.max_slice <= kBatchMaxStrikes && v.n_slices <= kPriceItems && n_units < kSplitUnits) {
    // fused path: one launch
    if (fd) {
      const size_t cb = (size_t)B * sizeof(unsigned int);
      if (ctx->d_counters.cap < cb || ctx->counters_zeroed < cb) {
        DHJ_CUDA(ctx, ctx->d_counters.reserve(cb));
        DHJ_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, ctx->d_counters.cap, ctx->stream));
        ctx->counters_zeroed = ctx->d_counters.cap;
      }
    }
    LossBatchArgs a;
    a.x = (const double*)ctx->d_x.p; a.market_index = d_index; a.S0 = (const double*)mk->d_S0.p;
    a.market = (const double*)mk->d_price.p; a.fd = fd; a.h = h; a.n_units = n_units;
    a.f_all = (double*)ctx->d_f.p; a.fg = fd ? (double*)ctx->d_fg.p : nullptr;
    a.counters = fd ? (unsigned int*)ctx->d_counters.p : nullptr;
    // whole units per block batch (pick_units_per_batch: waves x block time)
    a.units_per_batch = pick_units_per_batch(n_units, v.n_slices, kPriceItems,
                                             (long long)ctx->sm_count * ctx->loss_blocks_per_sm);
    const long long batches = (n_units + a.units_per_batch - 1) / a.units_per_batch;
    k_loss_batch<<<(int)std::max<long long>(1, std::min(batches, 2147483647LL)), kBatchThreads, 0, ctx->stream>>>(v, a);
    DHJ_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
  } else {
    // general path: stencil points -> pricing kernel (batch or dense by slice size) -> reduction
    DHJ_CUDA(ctx, ctx->d_xv.reserve((size_t)n_units * kNumParams * sizeof(double)));
    DHJ_CUDA(ctx, ctx->d_idx.reserve((size_t)n_units * sizeof(int)));
    DHJ_CUDA(ctx, ctx->d_prices.reserve((size_t)n_units * mk->M * sizeof(double)));
    const unsigned gb = (unsigned)((n_units + 127) / 128);
    k_fd_expand<<<gb, 128, 0, ctx->stream>>>((const double*)ctx->d_x.p, d_index, B, fd, h, (double*)ctx->d_xv.p,
                                             (int*)ctx->d_idx.p);
    DHJ_CUDA(ctx, cudaGetLastError());
    PriceArgs pa;
    pa.params = (const double*)ctx->d_xv.p; pa.S0 = (const double*)mk->d_S0.p; pa.s0_stride = 1;
    pa.row_index = (const int*)ctx->d_idx.p; pa.P = n_units; pa.transform = 0; pa.out = (double*)ctx->d_prices.p;
    ctx->launches++;
    rc = launch_price(ctx, v, pa, mk->book.max_slice, ctx->stream);
    if (rc) return rc;
    k_loss_reduce<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(
        (const double*)ctx->d_prices.p, (const double*)ctx->d_xv.p, (const int*)ctx->d_idx.p,
        (const double*)mk->d_price.p, v.pos, mk->M, B, fd, h, (const double*)ctx->d_x.p, (double*)ctx->d_f.p,
        fd ? (double*)ctx->d_fg.p : nullptr);
    DHJ_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
  }
  const bool want_all = fd && out_f_all;
  DHJ_CUDA(ctx, ctx->h_res.reserve(res_bytes * (want_all ? 2 : 1)));
  DHJ_CUDA(ctx, cudaMemcpyAsync(ctx->h_res.p, fd ? ctx->d_fg.p : ctx->d_f.p, res_bytes, cudaMemcpyDeviceToHost,
                                ctx->stream));
  if (want_all)
    DHJ_CUDA(ctx, cudaMemcpyAsync((unsigned char*)ctx->h_res.p + res_bytes, ctx->d_f.p, res_bytes,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  DHJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const double* res = (const double*)ctx->h_res.p;
  if (want_all) host_copy(out_f_all, (const unsigned char*)ctx->h_res.p + res_bytes, res_bytes);
  if (!fd) {
    host_copy(out_f, res, res_bytes);
  } else {
#pragma omp parallel for schedule(static) num_threads(copy_threads()) if (B >= 4096)
    for (int64_t c = 0; c < B; ++c) {
      out_f[c] = res[c * kFdPoints];
      memcpy(out_g + c * kNumParams, res + c * kFdPoints + 1, kNumParams * sizeof(double));
    }
  }
  return DHJ_OK;
}

int dhj_loss_batch(dhj_ctx* ctx, const dhj_market* market, const double* x, const int32_t* market_index,
                   int64_t B, double* out_loss) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  return run_loss(ctx, market, x, market_index, B, 0, 0.0, out_loss, nullptr, nullptr);
}

int dhj_loss_fd(dhj_ctx* ctx, const dhj_market* market, const double* x, const int32_t* market_index,
                int64_t C, double h, double* out_f, double* out_g, double* out_f_all) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!(h > 0.0)) return fail(ctx, DHJ_ERR_ARG, "h must be > 0");
  return run_loss(ctx, market, x, market_index, C, 1, h, out_f, out_g, out_f_all);
}

// The whole lock-step loop of many simultaneous calibrations in one call: ask -> one loss / forward-difference
// launch over every state that waits for an evaluation -> tell, until every optimiser has stopped.  (The Python
// host used to drive the three steps itself; with several pipelines per GPU their interpreter work serialised on
// the GIL — in here a pipeline never touches it.)
int dhj_lbfgs_minimize_fd(dhj_lbfgs* opt, dhj_ctx* ctx, const dhj_market* market, const int32_t* state_market,
                          int64_t n_states, double h, int64_t* rounds, int64_t* state_rounds, double* seconds) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!opt || !market) return fail(ctx, DHJ_ERR_ARG, "null optimiser or market");
  if (n_states < 0) return fail(ctx, DHJ_ERR_ARG, "n_states must be >= 0");
  if (!(h > 0.0)) return fail(ctx, DHJ_ERR_ARG, "h must be > 0");
  std::vector<int64_t> idx((size_t)n_states);
  std::vector<int32_t> mi((size_t)n_states);
  std::vector<double> x((size_t)n_states * kNumParams), f((size_t)n_states), g((size_t)n_states * kNumParams);
  int64_t n_rounds = 0, n_state_rounds = 0;
  double t_ask = 0.0, t_loss = 0.0, t_tell = 0.0;
  for (;;) {
    const double t0 = omp_get_wtime();
    int64_t na = 0;
    int rc = dhj_lbfgs_ask(opt, &na, idx.data(), x.data());
    if (rc) return fail(ctx, rc, "dhj_lbfgs_ask failed (%d)", rc);
    if (na > n_states) return fail(ctx, DHJ_ERR_ARG, "the optimiser holds more states (%lld) than n_states", (long long)na);
    const double t1 = omp_get_wtime();
    t_ask += t1 - t0;
    if (na == 0) break;
    if (state_market)
      for (int64_t a = 0; a < na; ++a) mi[a] = state_market[idx[a]];
    rc = run_loss(ctx, market, x.data(), state_market ? mi.data() : nullptr, na, 1, h, f.data(), g.data(), nullptr);
    if (rc) return rc;
    const double t2 = omp_get_wtime();
    rc = dhj_lbfgs_tell(opt, na, f.data(), g.data());
    if (rc) return fail(ctx, rc, "dhj_lbfgs_tell failed (%d)", rc);
    t_loss += t2 - t1;
    t_tell += omp_get_wtime() - t2;
    ++n_rounds;
    n_state_rounds += na;
  }
  if (rounds) *rounds = n_rounds;
  if (state_rounds) *state_rounds = n_state_rounds;
  if (seconds) { seconds[0] = t_ask; seconds[1] = t_loss; seconds[2] = t_tell; }
  return DHJ_OK;
}

int dhj_market_prices(dhj_ctx* ctx, const dhj_market* mk, const double* x, const int32_t* market_index,
                      int64_t B, double* out_prices) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  const int* d_index = nullptr;
  int rc = stage_x(ctx, mk, x, market_index, B, &d_index);
  if (rc) return rc;
  if (!out_prices) return fail(ctx, DHJ_ERR_ARG, "null output");
  if (B == 0) return DHJ_OK;
  const size_t ob = (size_t)B * mk->M * sizeof(double);
  DHJ_CUDA(ctx, ctx->d_prices.reserve(ob));
  DHJ_CUDA(ctx, ctx->h_res.reserve(ob));
  // without an index every x uses market 0: a zero table does that
  if (!d_index) {
    DHJ_CUDA(ctx, ctx->d_idx.reserve((size_t)B * sizeof(int)));
    DHJ_CUDA(ctx, cudaMemsetAsync(ctx->d_idx.p, 0, (size_t)B * sizeof(int), ctx->stream));
    d_index = (const int*)ctx->d_idx.p;
  }
  PriceArgs a;
  a.params = (const double*)ctx->d_x.p; a.S0 = (const double*)mk->d_S0.p; a.s0_stride = 1;
  a.row_index = d_index; a.P = B; a.transform = 1; a.out = (double*)ctx->d_prices.p;
  rc = launch_price(ctx, mk->view, a, mk->book.max_slice, ctx->stream);
  if (rc) return rc;
  DHJ_CUDA(ctx, cudaMemcpyAsync(ctx->h_res.p, ctx->d_prices.p, ob, cudaMemcpyDeviceToHost, ctx->stream));
  DHJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out_prices, ctx->h_res.p, ob);
  return DHJ_OK;
}

// ---- the remaining public methods of DoubleHeston --------------------------------------------
// small synchronous helper: host arrays in -> kernel -> host arrays out through the loss staging buffers
namespace {
struct Staged {
  dhj_ctx* ctx;
  size_t in_bytes = 0, out_bytes = 0;
  int begin(size_t in_b, size_t out_b) {
    in_bytes = BookHost::align8(in_b); out_bytes = out_b;
    DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
    DHJ_CUDA(ctx, ctx->h_x.reserve(in_bytes));
    DHJ_CUDA(ctx, ctx->d_x.reserve(in_bytes));
    DHJ_CUDA(ctx, ctx->d_prices.reserve(out_bytes));
    DHJ_CUDA(ctx, ctx->h_res.reserve(out_bytes));
    return DHJ_OK;
  }
  unsigned char* hin() { return (unsigned char*)ctx->h_x.p; }
  unsigned char* din() { return (unsigned char*)ctx->d_x.p; }
  int upload() {
    DHJ_CUDA(ctx, cudaMemcpyAsync(ctx->d_x.p, ctx->h_x.p, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    return DHJ_OK;
  }
  int download() {
    DHJ_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    DHJ_CUDA(ctx, cudaMemcpyAsync(ctx->h_res.p, ctx->d_prices.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DHJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DHJ_OK;
  }
};
}  // namespace

int dhj_cf(dhj_ctx* ctx, const double* params, double r, double q, double tau, const double* u, int32_t n,
           double* out_re, double* out_im) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!params || !u || !out_re || !out_im || n < 0) return fail(ctx, DHJ_ERR_ARG, "bad arguments");
  if (n == 0) return DHJ_OK;
  Staged st{ctx};
  const size_t pb = kNumParams * sizeof(double), ub = (size_t)n * sizeof(double);
  int rc = st.begin(pb + ub, 2 * ub);
  if (rc) return rc;
  memcpy(st.hin(), params, pb);
  memcpy(st.hin() + pb, u, ub);
  if ((rc = st.upload())) return rc;
  double* dout = (double*)ctx->d_prices.p;
  k_cf<<<(n + 127) / 128, 128, 0, ctx->stream>>>((const double*)st.din(), r, q, tau, (const double*)(st.din() + pb),
                                                  n, dout, dout + n);
  if ((rc = st.download())) return rc;
  memcpy(out_re, ctx->h_res.p, ub);
  memcpy(out_im, (unsigned char*)ctx->h_res.p + ub, ub);
  return DHJ_OK;
}

int dhj_cf_complex(dhj_ctx* ctx, const double* params, double r, double q, double tau, const double* u_re,
                   const double* u_im, int32_t n, double* out_re, double* out_im) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!params || !u_re || !u_im || !out_re || !out_im || n < 0) return fail(ctx, DHJ_ERR_ARG, "bad arguments");
  if (n == 0) return DHJ_OK;
  Staged st{ctx};
  const size_t pb = kNumParams * sizeof(double), ub = (size_t)n * sizeof(double);
  int rc = st.begin(pb + 2 * ub, 2 * ub);
  if (rc) return rc;
  memcpy(st.hin(), params, pb);
  memcpy(st.hin() + pb, u_re, ub);
  memcpy(st.hin() + pb + ub, u_im, ub);
  if ((rc = st.upload())) return rc;
  double* dout = (double*)ctx->d_prices.p;
  k_cf_complex<<<(n + 127) / 128, 128, 0, ctx->stream>>>((const double*)st.din(), r, q, tau,
                                                          (const double*)(st.din() + pb),
                                                          (const double*)(st.din() + pb + ub), n, dout, dout + n);
  if ((rc = st.download())) return rc;
  memcpy(out_re, ctx->h_res.p, ub);
  memcpy(out_im, (unsigned char*)ctx->h_res.p + ub, ub);
  return DHJ_OK;
}

int dhj_truncation_range(dhj_ctx* ctx, const double* params, int64_t P, const double* S0, int64_t s0_stride,
                         double r, const double* strike, const double* maturity, int32_t M, double L,
                         double* out_ab) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!params || !S0 || !strike || !maturity || !out_ab || P < 0 || M < 0) return fail(ctx, DHJ_ERR_ARG, "bad arguments");
  if (s0_stride != 0 && s0_stride != 1) return fail(ctx, DHJ_ERR_ARG, "s0_stride must be 0 or 1");
  if (P == 0 || M == 0) return DHJ_OK;
  if (P * (int64_t)M > (int64_t)1 << 28) return fail(ctx, DHJ_ERR_ARG, "P*M too large for this helper");
  Staged st{ctx};
  const size_t pb = (size_t)P * kNumParams * sizeof(double), sb = (s0_stride ? (size_t)P : 1) * sizeof(double);
  const size_t mb = (size_t)M * sizeof(double);
  int rc = st.begin(pb + sb + 2 * mb, (size_t)P * M * 2 * sizeof(double));
  if (rc) return rc;
  memcpy(st.hin(), params, pb);
  memcpy(st.hin() + pb, S0, sb);
  memcpy(st.hin() + pb + sb, strike, mb);
  memcpy(st.hin() + pb + sb + mb, maturity, mb);
  if ((rc = st.upload())) return rc;
  const long long n = P * (long long)M;
  k_truncation_range<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      (const double*)st.din(), P, (const double*)(st.din() + pb), s0_stride, (const double*)(st.din() + pb + sb),
      (const double*)(st.din() + pb + sb + mb), M, r, L, (double*)ctx->d_prices.p);
  if ((rc = st.download())) return rc;
  memcpy(out_ab, ctx->h_res.p, st.out_bytes);
  return DHJ_OK;
}

int dhj_chi_psi(dhj_ctx* ctx, const int32_t* k, int32_t n, double c, double d, double a, double b,
                double* out_chi, double* out_psi) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (!k || !out_chi || !out_psi || n < 0) return fail(ctx, DHJ_ERR_ARG, "bad arguments");
  if (n == 0) return DHJ_OK;
  Staged st{ctx};
  const size_t kb = (size_t)n * sizeof(int), ob = (size_t)n * sizeof(double);
  int rc = st.begin(kb, 2 * ob);
  if (rc) return rc;
  memcpy(st.hin(), k, kb);
  if ((rc = st.upload())) return rc;
  double* dout = (double*)ctx->d_prices.p;
  k_chi_psi<<<(n + 127) / 128, 128, 0, ctx->stream>>>((const int*)st.din(), n, c, d, a, b, dout, dout + n);
  if ((rc = st.download())) return rc;
  memcpy(out_chi, ctx->h_res.p, ob);
  memcpy(out_psi, (unsigned char*)ctx->h_res.p + ob, ob);
  return DHJ_OK;
}

}  // extern "C"

// ---- synthetic dataset sweep (counter stream; dhj_generate.cuh) ------------------------------------
namespace {

struct GenConfig {
  GenArgs g;
  int nK, nT, N;
  double r, L;
  const double* strikes_rel;
  const double* maturities;
};

int check_generate(dhj_ctx* ctx, int64_t first, int64_t n, int32_t path_len, const double* lo, const double* hi,
                   const double* strikes_rel, int32_t nK, const double* maturities, int32_t nT, int32_t N) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (first < 0 || n < 0) return fail(ctx, DHJ_ERR_ARG, "first and n must be >= 0");
  if (path_len < 1) return fail(ctx, DHJ_ERR_ARG, "path_len must be >= 1 (got %d)", path_len);
  if (!lo || !hi || !strikes_rel || !maturities) return fail(ctx, DHJ_ERR_ARG, "null table");
  if (nK < 1 || nT < 1 || (int64_t)nK * nT > kGenMaxOptions)
    return fail(ctx, DHJ_ERR_ARG, "the generator grid must have 1..%d options (got %d x %d)", kGenMaxOptions, nT, nK);
  if (N < 1) return fail(ctx, DHJ_ERR_ARG, "N must be >= 1 (got %d)", N);
  return DHJ_OK;
}

// enqueue draws -> prices -> market/loss for samples [first, first + n) on `st`; all pointers are device memory
int enqueue_generate(dhj_ctx* ctx, const GenConfig& c, int64_t first, int64_t n, double* d_params, double* d_spots,
                     double* d_model, double* d_market, double* d_loss, cudaStream_t st) {
  if (n == 0) return DHJ_OK;
  GenArgs g = c.g;
  g.first = first; g.n = n;
  const long long q_first = first / g.path_len, q_last = (first + n - 1) / g.path_len;
  const long long paths = q_last - q_first + 1;
  k_gen_draws<<<(unsigned)((paths + kGenWarps - 1) / kGenWarps), 32 * kGenWarps, 0, st>>>(g, d_params, d_spots);
  DHJ_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  SliceView v;
  int rc = grid_view(ctx, c.strikes_rel, c.nK, c.maturities, c.nT, /*scale_by_spot=*/1, /*is_call=*/1, c.N, c.r, 0.0,
                     c.L, st, &v);
  if (rc) return rc;
  PriceArgs a;
  a.params = d_params; a.S0 = d_spots; a.s0_stride = 1; a.row_index = nullptr; a.P = n; a.transform = 0;
  a.out = d_model;
  rc = launch_price(ctx, v, a, c.nK, st);
  if (rc) return rc;
  rc = book_used(ctx, st);
  if (rc) return rc;
  if (d_market && d_loss) {
    const int M = c.nK * c.nT;
    const size_t smem = (size_t)kGenMarketSamples * (M | 1) * sizeof(double);
    k_gen_market<<<(unsigned)((n + kGenMarketSamples - 1) / kGenMarketSamples), kGenMarketSamples, smem, st>>>(
        g.seed, first, n, M, g.noise_sd, d_model, d_market, d_loss);
    DHJ_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
  }
  return DHJ_OK;
}

void fill_gen_config(GenConfig* c, uint64_t seed, int32_t path_len, const double* lo, const double* hi,
                     double persistence, double spot0, double ret_mean, double ret_sd, double noise_sd,
                     const double* strikes_rel, int32_t nK, const double* maturities, int32_t nT, double r, int32_t N,
                     double L) {
  c->g.seed = seed; c->g.first = 0; c->g.n = 0; c->g.path_len = path_len;
  for (int j = 0; j < kNumParams; ++j) { c->g.lo[j] = lo[j]; c->g.range[j] = hi[j] - lo[j]; }
  c->g.persistence = persistence; c->g.spot0 = spot0; c->g.ret_mean = ret_mean; c->g.ret_sd = ret_sd;
  c->g.noise_sd = noise_sd;
  c->nK = nK; c->nT = nT; c->N = N; c->r = r; c->L = L; c->strikes_rel = strikes_rel; c->maturities = maturities;
}

}  // namespace

extern "C" {

int dhj_generate_dev(dhj_ctx* ctx, uint64_t seed, int64_t first, int64_t n, int32_t path_len, const double* lo,
                     const double* hi, double persistence, double spot0, double ret_mean, double ret_sd,
                     double noise_sd, const double* strikes_rel, int32_t nK, const double* maturities, int32_t nT,
                     double r, int32_t N, double L, double* d_params, double* d_spots, double* d_model,
                     double* d_market, double* d_loss, void* stream) {
  int rc = check_generate(ctx, first, n, path_len, lo, hi, strikes_rel, nK, maturities, nT, N);
  if (rc) return rc;
  if (!d_params || !d_spots || !d_model) return fail(ctx, DHJ_ERR_ARG, "null output (params, spots, model are required)");
  if ((d_market == nullptr) != (d_loss == nullptr))
    return fail(ctx, DHJ_ERR_ARG, "market and loss outputs come together (both or neither)");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  GenConfig c;
  fill_gen_config(&c, seed, path_len, lo, hi, persistence, spot0, ret_mean, ret_sd, noise_sd, strikes_rel, nK,
                  maturities, nT, r, N, L);
  return enqueue_generate(ctx, c, first, n, d_params, d_spots, d_model, d_market, d_loss, (cudaStream_t)stream);
}

int dhj_generate(dhj_ctx* ctx, uint64_t seed, int64_t first, int64_t n, int32_t path_len, const double* lo,
                 const double* hi, double persistence, double spot0, double ret_mean, double ret_sd, double noise_sd,
                 const double* strikes_rel, int32_t nK, const double* maturities, int32_t nT, double r, int32_t N,
                 double L, double* params, double* spots, double* model, double* market, double* loss) {
  int rc = check_generate(ctx, first, n, path_len, lo, hi, strikes_rel, nK, maturities, nT, N);
  if (rc) return rc;
  if (!params || !spots || !model || !market || !loss) return fail(ctx, DHJ_ERR_ARG, "null output array");
  if (n == 0) return DHJ_OK;
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  GenConfig c;
  fill_gen_config(&c, seed, path_len, lo, hi, persistence, spot0, ret_mean, ret_sd, noise_sd, strikes_rel, nK,
                  maturities, nT, r, N, L);
  const int M = nK * nT;
  const bool pinned = is_pinned_host(params) && is_pinned_host(spots) && is_pinned_host(model) &&
                      is_pinned_host(market) && is_pinned_host(loss);
  // chunks of whole paths where possible (a chunk that starts inside a path re-draws the path's head)
  int64_t chunk = 131072;
  if (path_len < chunk) chunk -= chunk % path_len;
  for (int i = 0; i < kSlots; ++i) {
    Slot& sl = ctx->slots[i];
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    sl.pending = false; sl.user_out = nullptr; sl.out_bytes = 0; sl.pieces.clear();
  }
  int slot_i = 0;
  for (int64_t lo_i = 0; lo_i < n; lo_i += chunk, slot_i = (slot_i + 1) % kSlots) {
    const int64_t cnt = std::min(chunk, n - lo_i);
    Slot& sl = ctx->slots[slot_i];
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    sl.drain_pieces(host_copy);
    // device layout of a chunk: params | spots | model | market | loss
    const size_t pb = (size_t)cnt * kNumParams * sizeof(double), sb = (size_t)cnt * sizeof(double);
    const size_t mb = (size_t)cnt * M * sizeof(double);
    const size_t total = pb + sb + 2 * mb + sb;
    DHJ_CUDA(ctx, sl.d_out.reserve(total));
    unsigned char* d = (unsigned char*)sl.d_out.p;
    double* d_params = (double*)d; double* d_spots = (double*)(d + pb); double* d_model = (double*)(d + pb + sb);
    double* d_market = (double*)(d + pb + sb + mb); double* d_loss = (double*)(d + pb + sb + 2 * mb);
    rc = enqueue_generate(ctx, c, first + lo_i, cnt, d_params, d_spots, d_model, d_market, d_loss, sl.stream);
    if (rc) return rc;
    void* dst[5] = {params + lo_i * kNumParams, spots + lo_i, model + lo_i * M, market + lo_i * M, loss + lo_i};
    const size_t off[5] = {0, pb, pb + sb, pb + sb + mb, pb + sb + 2 * mb};
    const size_t len[5] = {pb, sb, mb, mb, sb};
    if (pinned) {
      for (int k = 0; k < 5; ++k)
        DHJ_CUDA(ctx, cudaMemcpyAsync(dst[k], d + off[k], len[k], cudaMemcpyDeviceToHost, sl.stream));
    } else {
      DHJ_CUDA(ctx, sl.h_out.reserve(total));
      DHJ_CUDA(ctx, cudaMemcpyAsync(sl.h_out.p, d, total, cudaMemcpyDeviceToHost, sl.stream));
      for (int k = 0; k < 5; ++k) sl.pieces.push_back({dst[k], off[k], len[k]});
    }
    DHJ_CUDA(ctx, cudaEventRecord(sl.done, sl.stream));
  }
  for (int i = 0; i < kSlots; ++i) {
    Slot& sl = ctx->slots[i];
    DHJ_CUDA(ctx, cudaEventSynchronize(sl.done));
    sl.drain_pieces(host_copy);
  }
  return DHJ_OK;
}

}  // extern "C"

extern "C" {

// ---- checked build ---------------------------------------------------------------------------------
int dhj_debug_checks(dhj_ctx* ctx, int32_t* enabled, uint64_t* counts) {
  if (!ctx || !enabled || !counts) return fail(ctx, DHJ_ERR_ARG, "null argument");
  for (int i = 0; i < kChkCodes; ++i) counts[i] = 0;
#if defined(DHJ_CHECKED)
  *enabled = 1;
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  DHJ_CUDA(ctx, cudaDeviceSynchronize());
  unsigned long long host[kChkCodes];
  DHJ_CUDA(ctx, cudaMemcpyFromSymbol(host, g_check_fail, sizeof(host)));
  for (int i = 0; i < kChkCodes; ++i) counts[i] = host[i];
#else
  *enabled = 0;
#endif
  return DHJ_OK;
}

// ---- measurement -----------------------------------------------------------------------------
int dhj_fp64_peak(dhj_ctx* ctx, int32_t iters, double* tflops, double* milliseconds) {
  if (!ctx) return fail(nullptr, DHJ_ERR_ARG, "null context");
  if (iters < 1 || !tflops) return fail(ctx, DHJ_ERR_ARG, "bad arguments");
  DHJ_CUDA(ctx, cudaSetDevice(ctx->device));
  const int threads = 256, blocks = ctx->sm_count * 8;
  DHJ_CUDA(ctx, ctx->d_peak.reserve((size_t)threads * blocks * sizeof(double)));
  cudaEvent_t e0, e1;
  DHJ_CUDA(ctx, cudaEventCreate(&e0));
  DHJ_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 8; ++rep) {           // two operand forms (dhj_kernels.cuh), 4 repetitions each, the first a warm-up
    DHJ_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (rep < 4) k_fp64_peak<false><<<blocks, threads, 0, ctx->stream>>>((double*)ctx->d_peak.p, iters, 0.999999, 1e-7);
    else k_fp64_peak<true><<<blocks, threads, 0, ctx->stream>>>((double*)ctx->d_peak.p, iters, 0.999999, 1e-7);
    DHJ_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    DHJ_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    DHJ_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    DHJ_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep % 4 > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flop = 2.0 * (double)threads * blocks * (double)iters * kPeakChains * kPeakUnroll;
  *tflops = flop / ((double)best * 1e-3) / 1e12;
  if (milliseconds) *milliseconds = best;
  return DHJ_OK;
}

}  // extern "C"
