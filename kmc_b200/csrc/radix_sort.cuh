// kmc_b200 — 8-bit radix sort of packed k-mer records (replaces the reference's sort_func:
// RADULS kmc_core/raduls_impl.h:546-776 + first_dispatch.h:98-435, or radix.h:845-855 + small_sort.h).
//
// Contract kept (raduls.h:19-20, kb_sorter.h:757-780): ascending order on bytes key_bytes-1..0 of the
// record image.  Records are pure keys (no payload), so the sorted array is unique and any stable or
// unstable correct sort is bit-identical to RADULS' output.
//
// B200 design: one kernel launch per 8-bit digit, each pass exactly one read + one write of every
// record ("onesweep": chained scan with decoupled look-back gives every tile its global bucket
// offsets inside the same kernel).  Per pass and CTA:
//   * tiles are fetched global->shared with TMA 1-D bulk copies (cp.async.bulk + mbarrier),
//     double buffered: tile i+1 is in flight while tile i is ranked;
//   * ranks come from warp-private digit histograms in shared memory driven by match.any + popc
//     (warp-shuffle broadcast of the bucket cursor) - no atomics on the ranking path;
//   * records are regrouped by digit in shared memory and leave as digit-contiguous runs
//     (TILE/256 records = 128-256 B on average), so global stores are fully coalesced;
//   * the histogram of the NEXT digit is accumulated on the fly (the multiset of records does not
//     depend on their order), so no pass ever re-reads the data just to count:  traffic = 2*N*W / pass.
#pragma once
#include "common.cuh"

namespace kmcb {

struct SortPass {
	const void* in;          // N records
	void* out;               // N records
	uint64_t n;
	uint32_t n_tiles;
	uint32_t byte;           // digit = byte `byte` of the record image
	int32_t next_byte;       // digit whose histogram is accumulated for the following pass, -1: none
	const uint64_t* hist;    // [256] histogram of `byte` over all N records
	uint64_t* hist_next;     // [256] zero-initialised
	uint64_t* desc;          // [n_tiles][256] look-back descriptors (epoch-tagged, see common.cuh)
	uint32_t epoch;          // unique per launch
	uint32_t* tile_counter;  // zero-initialised
	const uint32_t* run_flag; // nullptr: always run; else run only when (*run_flag & 1): the hybrid MSD path gave up (msd_sort.cuh)
};

template <int WORDS> struct SortCfg;
template <> struct SortCfg<1> { static constexpr int kThreads = 512, kKpt = 8, kMinBlocks = 2; };
template <> struct SortCfg<2> { static constexpr int kThreads = 512, kKpt = 4, kMinBlocks = 2; };
template <> struct SortCfg<3> { static constexpr int kThreads = 512, kKpt = 3, kMinBlocks = 2; };
template <> struct SortCfg<4> { static constexpr int kThreads = 512, kKpt = 2, kMinBlocks = 2; };

template <int WORDS>
struct SortSmem {
	static constexpr int kThreads = SortCfg<WORDS>::kThreads;
	static constexpr int kWarps = kThreads / 32;
	static constexpr int kKpt = SortCfg<WORDS>::kKpt;
	static constexpr int kTile = kThreads * kKpt;
	static constexpr int kRecBytes = 8 * WORDS;
	static constexpr int kBufBytes = kTile * kRecBytes;
	// byte offsets inside dynamic shared memory
	static constexpr int oBuf = 0;
	static constexpr int oWhist = 2 * kBufBytes;              // u32 [kWarps][256]
	static constexpr int oTileExcl = oWhist + kWarps * 1024;  // u32 [256]
	static constexpr int oGoff = oTileExcl + 1024;            // u64 [256]
	static constexpr int oNhist = oGoff + 2048;               // u32 [256]
	static constexpr int oWarpTot = oNhist + 1024;            // u32 [32]
	static constexpr int oMbar = oWarpTot + 128;              // u64 [2]
	static constexpr int oTileId = oMbar + 16;                // u32 [2]
	static constexpr int kBytes = oTileId + 16;
};

// exclusive scan of one value per thread over the first 256 threads (8 warps); all threads must call.
__device__ __forceinline__ uint64_t block_excl_scan_256(uint64_t v, uint64_t* warp_tot /* smem [8] */, uint64_t* total)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint64_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (uint32_t)o) inc += t;
	}
	if (warp < 8 && lane == 31) warp_tot[warp] = inc;
	__syncthreads();
	uint64_t base = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < 8; ++w) {
		uint64_t t = warp_tot[w];
		if ((uint32_t)w < warp) base += t;
		tot += t;
	}
	if (total) *total = tot;
	__syncthreads();
	return base + inc - v;
}

template <int WORDS>
__global__ void __launch_bounds__(SortCfg<WORDS>::kThreads, SortCfg<WORDS>::kMinBlocks) radix_pass_kernel(const SortPass p)
{
	using S = SortSmem<WORDS>;
	using R = Rec<WORDS>;
	constexpr int THREADS = S::kThreads, WARPS = S::kWarps, KPT = S::kKpt, TILE = S::kTile;
	extern __shared__ __align__(128) uint8_t smem[];
	uint32_t* whist = reinterpret_cast<uint32_t*>(smem + S::oWhist);
	uint32_t* tile_excl = reinterpret_cast<uint32_t*>(smem + S::oTileExcl);
	uint64_t* goff = reinterpret_cast<uint64_t*>(smem + S::oGoff);
	uint32_t* nhist = reinterpret_cast<uint32_t*>(smem + S::oNhist);
	uint64_t* warp_tot = reinterpret_cast<uint64_t*>(smem + S::oWarpTot);
	uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + S::oMbar);
	volatile uint32_t* s_tile = reinterpret_cast<volatile uint32_t*>(smem + S::oTileId);

	if (p.run_flag && !(*p.run_flag & 1u)) return;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const R* __restrict__ gin = reinterpret_cast<const R*>(p.in);
	R* __restrict__ gout = reinterpret_cast<R*>(p.out);
	uint64_t* desc = p.desc;

	if (tid < 256) nhist[tid] = 0;
	if (tid == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		fence_mbar_init();
	}
	// bucket bases of this pass: exclusive scan of the digit histogram (thread d <-> digit d)
	uint64_t bucket_base = block_excl_scan_256(tid < 256 ? p.hist[tid] : 0, warp_tot, nullptr);

	auto issue_load = [&](uint32_t tile, int b) {
		// called by thread 0 only
		const uint64_t first = (uint64_t)tile * TILE;
		const uint64_t rem = p.n - first;
		const uint32_t valid = rem < (uint64_t)TILE ? (uint32_t)rem : (uint32_t)TILE;
		const uint32_t bytes = valid * S::kRecBytes;
		if ((bytes & 15u) == 0) {       // TMA path; an odd-sized last tile is fetched by all threads when it is consumed
			fence_proxy_async();
			mbar_arrive_expect_tx(&mbar[b], bytes);
			bulk_g2s(smem + S::oBuf + b * S::kBufBytes, gin + first, bytes, &mbar[b]);
		}
	};

	if (tid == 0) {
		uint32_t t = atomicAdd(p.tile_counter, 1u);
		s_tile[0] = t;
		if (t < p.n_tiles) issue_load(t, 0);
	}
	__syncthreads();

	uint32_t phase0 = 0, phase1 = 0, cnt_real = 0;
	int cur = 0;
	while (true) {
		const uint32_t tile = s_tile[cur];
		if (tile >= p.n_tiles) break;
		if (tid == 0) {   // claim and prefetch the next tile of this CTA
			uint32_t t = atomicAdd(p.tile_counter, 1u);
			s_tile[cur ^ 1] = t;
			if (t < p.n_tiles) issue_load(t, cur ^ 1);
		}
		const uint64_t first = (uint64_t)tile * TILE;
		const uint64_t rem = p.n - first;
		const uint32_t valid = rem < (uint64_t)TILE ? (uint32_t)rem : (uint32_t)TILE;
		R* buf = reinterpret_cast<R*>(smem + S::oBuf + cur * S::kBufBytes);

		// zero the warp-private histograms
#pragma unroll
		for (int i = tid; i < WARPS * 256; i += THREADS) whist[i] = 0;

		if (((valid * S::kRecBytes) & 15u) == 0) {
			if (cur == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
			else { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
		} else {
			for (uint32_t i = tid; i < valid; i += THREADS) buf[i] = gin[first + i];
		}
		__syncthreads();

		// ---- phase 1: every record into registers + warp-private digit counts (shared-memory atomics, conflicts only inside a warp).
		// Warp w owns records [w*32*KPT, (w+1)*32*KPT) of the tile, round r covers 32 consecutive ones.
		R key[KPT];
		uint32_t* wh = whist + warp * 256;
#pragma unroll
		for (int r = 0; r < KPT; ++r) {
			const uint32_t idx = warp * (32 * KPT) + r * 32 + lane;
			if (idx < valid) key[r] = buf[idx];
			else {
#pragma unroll
				for (int j = 0; j < WORDS; ++j) key[r].w[j] = ~0ull;     // padding sorts to the very end of digit 255
			}
			atomicAdd(&wh[rec_byte<WORDS>(key[r], p.byte)], 1u);
		}
		__syncthreads();

		// ---- phase 2, digit d (thread d): offsets of every warp inside the digit's run, tile count; the aggregate is
		// published BEFORE the expensive ranking so that later tiles almost never wait in their look-back
		uint32_t cnt = 0;
		if (tid < 256) {
#pragma unroll
			for (int w = 0; w < WARPS; ++w) {
				uint32_t t = whist[w * 256 + tid];
				whist[w * 256 + tid] = cnt;
				cnt += t;
			}
			uint32_t real = cnt;
			if (tid == 255) real -= (TILE - valid);       // padding records are not published
			st_relaxed(desc + (uint64_t)tile * 256 + tid, desc_pack(tile == 0 ? kDescPrefix : kDescAggregate, p.epoch, real));
			cnt_real = real;
		}
		const uint64_t texcl = block_excl_scan_256(cnt, warp_tot, nullptr);
		if (tid < 256) tile_excl[tid] = (uint32_t)texcl;
		__syncthreads();

		// ---- phase 3: stable ranks from match.any + the warp's running bucket cursor; regroup by digit in shared memory
		// (the tile buffer is dead: every record is in registers).  All match.any are issued first (independent, their
		// latency overlaps); only the short cursor update is a serial chain over the rounds.
		uint32_t peers[KPT];
#pragma unroll
		for (int r = 0; r < KPT; ++r) peers[r] = __match_any_sync(0xffffffffu, rec_byte<WORDS>(key[r], p.byte));
#pragma unroll
		for (int r = 0; r < KPT; ++r) {
			const uint32_t d = rec_byte<WORDS>(key[r], p.byte);
			const uint32_t m = peers[r];
			const uint32_t below = __popc(m & lanemask_lt());
			const int leader = __ffs(m) - 1;
			uint32_t old = 0;
			if ((int)lane == leader) {
				old = wh[d];
				wh[d] = old + __popc(m);
			}
			old = __shfl_sync(0xffffffffu, old, leader);
			buf[tile_excl[d] + old + below] = key[r];
			__syncwarp();
		}

		// ---- phase 4: chained scan over tiles (decoupled look-back), one digit per thread
		if (tid < 256) {
			const uint64_t excl = tile == 0 ? 0 : lookback_resolve(desc + tid, 256, tile, (uint64_t)cnt_real, p.epoch);
			goff[tid] = bucket_base + excl - texcl;    // global index of tile-sorted position q is goff[d] + q
		}
		__syncthreads();

		// ---- phase 5: digit-contiguous runs leave with coalesced stores; count the next digit on the way out
#pragma unroll
		for (int i = 0; i < KPT; ++i) {
			const uint32_t q = i * THREADS + tid;
			if (q < valid) {
				const R k = buf[q];
				const uint32_t d = rec_byte<WORDS>(k, p.byte);
				gout[goff[d] + q] = k;
				if (p.next_byte >= 0) atomicAdd(&nhist[rec_byte<WORDS>(k, (uint32_t)p.next_byte)], 1u);
			}
		}
		fence_proxy_async();     // generic-proxy accesses of this buffer are ordered before the TMA refill
		__syncthreads();
		cur ^= 1;
	}

	if (p.next_byte >= 0 && tid < 256) {
		const uint32_t c = nhist[tid];
		if (c) atomicAdd(reinterpret_cast<unsigned long long*>(p.hist_next) + tid, (unsigned long long)c);
	}
}

// histogram of the first digit over N records, in front of the LSD passes (run_flag as in SortPass)
template <int WORDS>
__global__ void __launch_bounds__(512) digit_histogram_kernel(const void* in, uint64_t n, uint32_t byte, uint64_t* hist, const uint32_t* run_flag)
{
	__shared__ uint32_t sh[256];
	if (run_flag && !(*run_flag & 1u)) return;
	const Rec<WORDS>* __restrict__ g = reinterpret_cast<const Rec<WORDS>*>(in);
	if (threadIdx.x < 256) sh[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	// one CTA never sees 2^32 records of one digit (n / gridDim.x < 2^32)
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		atomicAdd(&sh[rec_byte<WORDS>(g[i], byte)], 1u);
	__syncthreads();
	if (threadIdx.x < 256) { uint32_t c = sh[threadIdx.x]; if (c) atomicAdd((unsigned long long*)hist + threadIdx.x, (unsigned long long)c); }
}

}  // namespace kmcb
