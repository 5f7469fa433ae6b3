// kmc_b200 — leaves of the hybrid MSD path: COUNT WITHOUT SORTING THE DUPLICATES, one WARP per leaf.
//
// What stage 2 needs from a leaf bucket (records that share their top 8+b2 bits, ~1 K records) is the sorted list of its
// DISTINCT k-mers with their multiplicities (CompactKmers, kmc_core/kb_sorter.h:1128-1281; for k % 32 != 0 the same result
// comes out of CompactKxmers :937-1122 + kxmer_set.h).  With sequencing coverage c a k-mer occurs ~c times, so sorting every
// copy (RADULS' lower levels and small sorts, raduls_impl.h:77-141,493-519) does ~c times the necessary work.  A leaf is
// therefore counted in an ORDER-PRESERVING table:
//   * slot = the next SLOT_BITS bits of the k-mer (monotone in the k-mer: slot order is key order, nothing is sorted);
//     the first copy claims the slot with one 64-bit atomicCAS, every other copy is one 32-bit atomicAdd on the count;
//   * a different k-mer that maps to a taken slot is noted and goes, in a dense second step, to a small open-addressing side
//     table; surviving side entries are ranked against the few entries of their own slot only;
//   * cutoffs / clamp / record bytes / lut[prefix]++ exactly as kb_sorter.h:1174-1203; one sweep over the slots gives every
//     surviving k-mer its position; records are written lane-dense (coalesced) into the leaf's region of a temporary buffer,
//     leaf_scan_kernel + leaf_gather_kernel pack the regions into the output.
// Everything is private to ONE WARP (its own table in shared memory, __syncwarp only): no CTA barrier, no inter-warp
// dependency, so the short latency-bound phases of one leaf overlap with those of ~30 other leaves per SM.
// A leaf larger than a round of the table (canonical k-mers crowd into the low prefixes: up to ~4.4x the mean) is counted in
// 2^e rounds over sub-ranges of its next e bits; a round whose tables overflow anyway is split in two on the next bit (binary
// descent, nothing has been emitted for it yet).  Only a leaf beyond kLwMaxLeaf records (or one k-mer-range that cannot be
// split any further) raises the device flag: the LSD passes + count_emit_kernel behind then redo the bin.
//
// Entry formats (64 bit, EMPTY = all ones):
//   WORDS == 1   [ key bits below the slot bits (<= 47) | count ]      the k-mer is rebuilt from leaf, slot and entry
//   WORDS >= 2   [ index of the first copy inside the leaf (32) | count (32) ]   equality is checked against that record
#pragma once
#include "common.cuh"
#include "expand.cuh"
#include "msd_sort.cuh"
#include "count.cuh"

namespace kmcb {

constexpr int kLwWarps = 4;                      // warps per CTA (independent of each other)
constexpr int kLwSide = 128;                     // side table (open addressing)
constexpr int kLwSideMax = 100;
constexpr int kLwRetry = 256;                    // records whose slot was taken, per round
constexpr uint32_t kLwMaxLeaf = 65534;           // records of a warp-counted leaf (u16 indices; count field >= 16 bits)
constexpr uint32_t kLwMaxSplit = 10;             // extra split bits a round may descend
constexpr uint64_t kLwEmpty = ~0ull;

struct LeafArgs {
	const void* recs;            // partitioned records
	const uint64_t* start;       // [n_leaves + 1]
	uint32_t n_leaves;
	uint32_t low_bits;           // bits below the partition digits
	uint32_t k, lut_prefix_len, cutoff_min, cutoff_max, counter_max, counter_bytes, suffix_bytes;
	uint8_t* tmp;                // leaf L writes its records, padded to a multiple of 8 bytes, at tmp + start[L] * pad
	uint32_t* leaf_emit;         // [n_leaves] emitted records
	uint64_t* lut;
	uint64_t* result;            // [0] n_unique [1] n_cutoff_min [2] n_cutoff_max
	uint32_t* ticket;
	uint32_t* flags;
};

template <int SLOT_BITS>
struct LwSmem {
	static constexpr int kSlots = 1 << SLOT_BITS;
	uint64_t main[kSlots];           // main table
	uint64_t skey[kLwSide];          // side table: the k-mer (WORDS == 1) or the index of its first copy
	uint32_t scnt[kLwSide];
	uint32_t extra[kSlots / 4];      // surviving side entries per slot, one byte each
	uint16_t gbase[kSlots / 4];      // output position of the first survivor of a group of 4 slots (written where side entries exist)
	uint16_t list[kLwRetry];         // insertion: records whose slot was taken; emission: survivors in output order
	uint8_t dense[kLwSide];          // surviving side entries
};

template <int WORDS>
__device__ __forceinline__ Rec<WORDS> lw_load(const Rec<WORDS>* p)
{
	Rec<WORDS> r;
	if (WORDS == 2) {
		const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
		r.w[0] = v.x; r.w[WORDS - 1] = v.y;
	} else if (WORDS == 4) {
		const ulonglong2 v0 = __ldg(reinterpret_cast<const ulonglong2*>(p)), v1 = __ldg(reinterpret_cast<const ulonglong2*>(p) + 1);
		r.w[0] = v0.x; r.w[1 % WORDS] = v0.y; r.w[2 % WORDS] = v1.x; r.w[3 % WORDS] = v1.y;
	} else {
#pragma unroll
		for (int i = 0; i < WORDS; ++i) r.w[i] = __ldg(reinterpret_cast<const unsigned long long*>(p) + i);
	}
	return r;
}

// 64 bits of the record starting at bit `pos` (multiple of 8, < 64 * WORDS)
template <int WORDS>
__device__ __forceinline__ uint64_t lw_extract64(const Rec<WORDS>& r, uint32_t pos)
{
	const uint32_t wi = pos >> 6, off = pos & 63u;
	uint64_t lo = r.w[0], hi = 0;
#pragma unroll
	for (int i = 1; i < WORDS; ++i) {
		if (wi == (uint32_t)i) lo = r.w[i];
		if (wi + 1 == (uint32_t)i) hi = r.w[i];
	}
	if (WORDS == 1 || wi + 1 >= (uint32_t)WORDS) hi = 0;
	uint64_t v = lo >> off;
	if (off) v |= hi << (64u - off);
	return v;
}

// word w of the emitted record: (k-p)/4 suffix bytes most significant first, then the counter least significant first
// (kb_sorter.h:1198-1201), as little-endian 64-bit words (byte 0 of the record = bits 0-7 of word 0)
template <int WORDS>
__device__ __forceinline__ uint64_t lw_out_word(const Rec<WORDS>& key, uint32_t value, uint32_t sb, uint32_t w)
{
	const int nkb = (int)sb - 8 * (int)w;          // suffix bytes still to go out from this word on
	if (nkb >= 8) return bswap64(lw_extract64<WORDS>(key, 8u * (uint32_t)(nkb - 8)));
	if (nkb > 0) return bswap64(key.w[0] << (8 * (8 - nkb))) | ((uint64_t)value << (8 * nkb));
	const int sh = -nkb;
	return sh < 4 ? (uint64_t)(value >> (8 * sh)) : 0ull;
}

template <int WORDS>
__device__ __forceinline__ uint32_t lw_hash(const Rec<WORDS>& r)
{
	uint64_t x = r.w[0];
#pragma unroll
	for (int i = 1; i < WORDS; ++i) x = (x ^ (x >> 29)) * 0xBF58476D1CE4E5B9ull + r.w[i];
	return (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> 40);
}

__device__ __forceinline__ uint32_t lw_bytesum(uint32_t x) { return (x * 0x01010101u) >> 24; }

template <int WORDS, int SLOT_BITS>
__global__ void __launch_bounds__(32 * kLwWarps, 7) leaf_warp_kernel(const LeafArgs a)
{
	using R = Rec<WORDS>;
	using SM = LwSmem<SLOT_BITS>;
	constexpr int SLOTS = SM::kSlots;
	constexpr uint32_t ROUND = SLOTS + SLOTS / 4;        // records a round is sized for
	constexpr int U = WORDS == 1 ? 4 : 2;                // records per lane and step
	constexpr uint32_t FULL = 0xffffffffu;
	extern __shared__ __align__(16) uint8_t lw_dsm[];
	SM& S = reinterpret_cast<SM*>(lw_dsm)[threadIdx.x >> 5];
	if (*a.flags & kMsdFlagFallback) return;
	const uint32_t lane = threadIdx.x & 31u, lt = lanemask_lt();
	const R* __restrict__ recs = reinterpret_cast<const R*>(a.recs);
	const uint32_t ob = a.suffix_bytes + a.counter_bytes;
	const uint32_t padw = (ob + 7) >> 3;                                   // temporary records: padw 64-bit words
	const uint32_t prefix_shift = 2u * (a.k - a.lut_prefix_len);
	const bool one_prefix = prefix_shift >= a.low_bits;                    // every k-mer of a leaf has the same LUT prefix
	const uint32_t span = a.cutoff_max - a.cutoff_min;                     // survivor <=> count - cutoff_min <= span (cutoff_max >= cutoff_min) ...
	const bool never = a.cutoff_max < a.cutoff_min;                        // ... unless nothing can survive
	uint64_t* const tmp64 = reinterpret_cast<uint64_t*>(a.tmp);
	uint32_t n_unique = 0, n_min = 0, n_max = 0;
	bool failed = false;

	uint32_t leaf = 0;
	if (lane == 0) leaf = atomicAdd(a.ticket, 1u);
	leaf = __shfl_sync(FULL, leaf, 0);
	while (leaf < a.n_leaves) {
		uint32_t next_t = 0;
		if (lane == 0) next_t = atomicAdd(a.ticket, 1u);                   // the next leaf: in flight while this one is counted
		const uint64_t lo = a.start[leaf];
		const uint32_t m = (uint32_t)min(a.start[leaf + 1] - lo, (uint64_t)0xffffffffu);
		uint32_t emit_base = 0;
		bool prefetched = false;
		if (m > kLwMaxLeaf) failed = true;
		else if (m > 0) {
			uint32_t e0 = 0;
			while ((m >> e0) > ROUND && e0 < 8 && e0 < a.low_bits) ++e0;
			uint32_t e = e0, r = 0;
			while (true) {
				// ================================================================ one round: the k-mers whose next e bits are r
				const uint32_t sub_shift = a.low_bits - e;
				const uint32_t slot_shift = sub_shift > (uint32_t)SLOT_BITS ? sub_shift - SLOT_BITS : 0;
				const uint32_t cb = WORDS == 1 ? min(64u - slot_shift, 32u) : 32u;              // bits of the count field
				const uint64_t rem_mask = (1ull << slot_shift) - 1ull;                           // slot_shift <= 47
				const uint32_t cmask = cb >= 32 ? 0xffffffffu : ((1u << cb) - 1u);
				const uint32_t emask = (1u << e) - 1u;
				// ---- clear
				{
					uint4* m4 = reinterpret_cast<uint4*>(S.main);
					const uint4 ev = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
					for (int i = 0; i < SLOTS * 8 / 16 / 32; ++i) m4[i * 32 + lane] = ev;
#pragma unroll
					for (int i = 0; i < kLwSide * 8 / 16 / 32; ++i) reinterpret_cast<uint4*>(S.skey)[i * 32 + lane] = ev;
					reinterpret_cast<uint4*>(S.scnt)[lane] = make_uint4(0, 0, 0, 0);              // 128 * 4 B
#pragma unroll
					for (int i = 0; i < SLOTS / 4 * 4 / 16 / 32; ++i) reinterpret_cast<uint4*>(S.extra)[i * 32 + lane] = make_uint4(0, 0, 0, 0);
					if (SLOTS / 4 * 4 / 16 < 32) { if (lane < SLOTS / 4 * 4 / 16) reinterpret_cast<uint4*>(S.extra)[lane] = make_uint4(0, 0, 0, 0); }
				}
				__syncwarp();
				// ---- insertion: first copy claims the slot, other copies add one, k-mers that find their slot taken are noted
				uint32_t n_retry = 0;
				bool ok = true;
				for (uint32_t j0 = 0; j0 < m; j0 += U * 32) {
					R key[U];
#pragma unroll
					for (int u = 0; u < U; ++u) {
						const uint32_t j = j0 + u * 32 + lane;
						if (j < m) key[u] = lw_load<WORDS>(recs + lo + j);
						else {
#pragma unroll
							for (int i = 0; i < WORDS; ++i) key[u].w[i] = 0;
						}
					}
					uint32_t slot[U];
					unsigned long long old[U];
					uint32_t live = 0;
#pragma unroll
					for (int u = 0; u < U; ++u) {              // all claims are issued before any result is looked at
						const uint32_t j = j0 + u * 32 + lane;
						bool in = j < m;
						if (e && rec_bits<WORDS>(key[u], sub_shift, emask) != r) in = false;      // another round's k-mer
						slot[u] = rec_bits<WORDS>(key[u], slot_shift, SLOTS - 1);
						old[u] = 0;
						if (in) {
							const unsigned long long ent = WORDS == 1 ? (((key[u].w[0] & rem_mask) << cb) | 1ull) : (((unsigned long long)j << 32) | 1ull);
							old[u] = atomicCAS(reinterpret_cast<unsigned long long*>(&S.main[slot[u]]), (unsigned long long)kLwEmpty, ent);
							live |= 1u << u;
						}
					}
#pragma unroll
					for (int u = 0; u < U; ++u) {
						bool coll = false;
						if (((live >> u) & 1u) && old[u] != kLwEmpty) {
							bool same;
							if (WORDS == 1) same = (old[u] >> cb) == (key[u].w[0] & rem_mask);
							else same = rec_equal<WORDS>(lw_load<WORDS>(recs + lo + (uint32_t)(old[u] >> 32)), key[u]);
							if (same) atomicAdd(reinterpret_cast<uint32_t*>(&S.main[slot[u]]), 1u);       // low word = count
							else coll = true;
						}
						const uint32_t cm = __ballot_sync(FULL, coll);
						if (cm) {
							const uint32_t q = n_retry + __popc(cm & lt);
							if (coll && q < (uint32_t)kLwRetry) S.list[q] = (uint16_t)(j0 + u * 32 + lane);
							n_retry += __popc(cm);
						}
					}
					if (n_retry > (uint32_t)kLwRetry) { ok = false; break; }
				}
				if (!prefetched) {        // the next leaf: towards L2 while this one is counted
					prefetched = true;
					const uint32_t nl = __shfl_sync(FULL, next_t, 0);
					if (nl < a.n_leaves) {
						const uint64_t nlo = a.start[nl];
						const uint32_t nm = (uint32_t)min(a.start[nl + 1] - nlo, (uint64_t)kLwMaxLeaf);
						for (uint32_t i = lane * (128 / (8 * WORDS)); i < nm; i += 32 * (128 / (8 * WORDS))) asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + nlo + i));
					}
				}
				__syncwarp();
				// ---- the noted records go to the side table
				uint32_t n_side = 0;
				if (ok && n_retry) {
					for (uint32_t q = lane; q < n_retry; q += 32) {
						const uint32_t j = S.list[q];
						const R kk = lw_load<WORDS>(recs + lo + j);
						uint32_t h = lw_hash<WORDS>(kk) & (kLwSide - 1);
						const unsigned long long mine = WORDS == 1 ? (unsigned long long)kk.w[0] : (unsigned long long)j;
						int probe = 0;
						for (; probe < kLwSide; ++probe) {
							const unsigned long long o2 = atomicCAS(reinterpret_cast<unsigned long long*>(&S.skey[h]), (unsigned long long)kLwEmpty, mine);
							bool hit = o2 == kLwEmpty;
							if (hit) ++n_side;
							else if (WORDS == 1) hit = o2 == mine;
							else hit = rec_equal<WORDS>(lw_load<WORDS>(recs + lo + (uint32_t)o2), kk);
							if (hit) { atomicAdd(&S.scnt[h], 1u); break; }
							h = (h + 1) & (kLwSide - 1);
						}
						if (probe == kLwSide) ok = false;
					}
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) n_side += __shfl_xor_sync(FULL, n_side, o);
					if (n_side > (uint32_t)kLwSideMax) ok = false;
					__syncwarp();
				}
				ok = __all_sync(FULL, ok);
				if (!ok) {        // this range does not fit: split it on the next bit (nothing of it has been emitted)
					if (e < a.low_bits && e < e0 + kLwMaxSplit) { ++e; r <<= 1; continue; }
					failed = true;
					break;
				}
				// the k-mer of a main / side entry
				const uint64_t key_hi = (WORDS > 1 || (slot_shift + SLOT_BITS) >= 64) ? 0ull
					: (((((uint64_t)leaf << a.low_bits) | ((uint64_t)r << sub_shift)) >> (slot_shift + SLOT_BITS)) << (slot_shift + SLOT_BITS));
				auto main_key = [&](uint32_t s, uint64_t ent) -> R {
					R kk;
					if (WORDS == 1) kk.w[0] = key_hi | ((uint64_t)s << slot_shift) | ((ent >> cb) & rem_mask);
					else kk = lw_load<WORDS>(recs + lo + (uint32_t)(ent >> 32));
					return kk;
				};
				auto side_key = [&](uint32_t h) -> R {
					R kk;
					if (WORDS == 1) kk.w[0] = S.skey[h];
					else kk = lw_load<WORDS>(recs + lo + (uint32_t)S.skey[h]);
					return kk;
				};
				auto survives = [&](uint32_t c) -> bool { return !never && (c - a.cutoff_min) <= span; };     // kb_sorter.h:1174-1191
				// ---- side entries: cutoffs, the survivors as a list + one byte per slot
				uint32_t n_dense = 0;
				if (n_side) {
#pragma unroll
					for (int i = 0; i < kLwSide / 32; ++i) {
						const uint32_t h = i * 32 + lane;
						const bool occ = S.skey[h] != kLwEmpty;
						const uint32_t c = S.scnt[h];
						const bool sv = occ && survives(c);
						n_min += occ && c < a.cutoff_min;
						n_max += occ && !sv && c >= a.cutoff_min;
						if (sv) {
							const uint32_t s = rec_bits<WORDS>(side_key(h), slot_shift, SLOTS - 1);
							atomicAdd(&S.extra[s >> 2], 1u << (8u * (s & 3u)));
						}
						const uint32_t sm = __ballot_sync(FULL, sv);
						if (sv) S.dense[n_dense + __popc(sm & lt)] = (uint8_t)h;
						n_dense += __popc(sm);
					}
					n_unique += lane == 0 ? n_side : 0;
					__syncwarp();
				}
				// ---- one sweep over the slots (4 consecutive slots per lane and step): cutoffs, positions, survivors listed in
				// output order; the list is emitted lane-dense whenever the next step might not fit
				uint32_t running = 0, listed_from = 0;          // survivors so far in this round; first position held by the list
				auto flush = [&](uint32_t upto) {                // emits the listed survivors [listed_from, upto)
					__syncwarp();
					for (uint32_t p = listed_from + lane; p < upto; p += 32) {
						const uint32_t v = S.list[p - listed_from];
						R kk; uint32_t c;
						if (v & 0x8000u) { const uint32_t h = v & 0x7fffu; kk = side_key(h); c = S.scnt[h]; }
						else { const uint64_t ent = S.main[v]; kk = main_key(v, ent); c = (uint32_t)ent & cmask; }
						const uint32_t value = c > a.counter_max ? a.counter_max : c;          // kb_sorter.h:1190
						uint64_t* dst = tmp64 + (lo + emit_base + p) * padw;
						for (uint32_t w = 0; w < padw; ++w) dst[w] = lw_out_word<WORDS>(kk, value, a.suffix_bytes, w);
						if (!one_prefix) atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + rec_prefix<WORDS>(kk, prefix_shift), 1ull);     // kb_sorter.h:1203
					}
					__syncwarp();
					listed_from = upto;
				};
				for (int it = 0; it < SLOTS / 128; ++it) {
					const uint32_t g = it * 32 + lane;           // group of 4 slots
					if (running - listed_from + 128u + n_dense > (uint32_t)kLwRetry) flush(running);
					const uint4 x0 = reinterpret_cast<const uint4*>(S.main)[2 * g], x1 = reinterpret_cast<const uint4*>(S.main)[2 * g + 1];
					const uint32_t xs = n_dense ? S.extra[g] : 0u;
					const uint64_t ent[4] = {((uint64_t)x0.y << 32) | x0.x, ((uint64_t)x0.w << 32) | x0.z, ((uint64_t)x1.y << 32) | x1.x, ((uint64_t)x1.w << 32) | x1.z};
					uint32_t nib = 0;
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const bool occ = ent[i] != kLwEmpty;
						const uint32_t c = (uint32_t)ent[i] & cmask;
						const bool sv = occ && survives(c);
						n_unique += occ;
						n_min += occ && c < a.cutoff_min;
						n_max += occ && !sv && c >= a.cutoff_min;
						nib |= sv ? (1u << i) : 0u;
					}
					const uint32_t c4 = __popc(nib) + lw_bytesum(xs);
					uint32_t inc = c4;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const uint32_t t = __shfl_up_sync(FULL, inc, o);
						if (lane >= (uint32_t)o) inc += t;
					}
					const uint32_t base = running + inc - c4;
					running += __shfl_sync(FULL, inc, 31);
					if (xs) S.gbase[g] = (uint16_t)base;
					uint32_t off = 0;
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const uint32_t xi = (xs >> (8 * i)) & 0xffu;
						if (((nib >> i) & 1u) && xi == 0) S.list[base + off - listed_from] = (uint16_t)(4 * g + i);      // slots with side entries: below
						off += ((nib >> i) & 1u) + xi;
					}
					if (__any_sync(FULL, xs != 0)) {
						// ---- slots that hold surviving side entries: every side entry ranks itself among the entries of its slot
						__syncwarp();
						for (uint32_t f = lane; f < n_dense; f += 32) {
							const uint32_t h = S.dense[f];
							const R kk = side_key(h);
							const uint32_t s = rec_bits<WORDS>(kk, slot_shift, SLOTS - 1);
							if ((s >> 7) != (uint32_t)it) continue;
							const uint32_t gg = s >> 2, ii = s & 3u;
							const uint32_t xg = S.extra[gg];
							uint32_t o2 = 0;
							bool main_sv = false;
							uint64_t main_ent = 0;
#pragma unroll
							for (int i = 0; i < 4; ++i) {
								const uint64_t en = S.main[4 * gg + i];
								const bool sv = en != kLwEmpty && survives((uint32_t)en & cmask);
								if ((uint32_t)i < ii) o2 += (sv ? 1u : 0u) + ((xg >> (8 * i)) & 0xffu);
								if ((uint32_t)i == ii) { main_sv = sv; main_ent = en; }
							}
							const uint32_t p0 = (uint32_t)S.gbase[gg] + o2;              // first position of the slot
							const uint32_t xi = (xg >> (8 * ii)) & 0xffu;
							uint32_t rank = 0, side_lt_main = 0;
							R mk;
							if (main_sv) { mk = main_key(s, main_ent); rank += rec_less<WORDS>(mk, kk) ? 1u : 0u; }
							bool first = true;                                            // the smallest side entry of the slot also places the main entry
							if (xi > 1) {
								for (uint32_t f2 = 0; f2 < n_dense; ++f2) {
									if (f2 == f) continue;
									const R k2 = side_key(S.dense[f2]);
									if (rec_bits<WORDS>(k2, slot_shift, SLOTS - 1) != s) continue;
									if (rec_less<WORDS>(k2, kk)) { ++rank; first = false; }
									if (main_sv && rec_less<WORDS>(k2, mk)) ++side_lt_main;
								}
							}
							S.list[p0 + rank - listed_from] = (uint16_t)(0x8000u | h);
							if (main_sv && first) {
								side_lt_main += rec_less<WORDS>(kk, mk) ? 1u : 0u;
								S.list[p0 + side_lt_main - listed_from] = (uint16_t)s;
							}
						}
					}
				}
				flush(running);
				emit_base += running;
				// ---- next round: back up from finished halves of a split, then one step to the right
				while (e > e0 && (r & 1u)) { r >>= 1; --e; }
				++r;
				if (e == e0 && r == (1u << e0)) break;
			}
		}
		if (lane == 0) {
			a.leaf_emit[leaf] = failed ? 0u : emit_base;
			if (one_prefix && emit_base && !failed)
				atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + (leaf >> (prefix_shift - a.low_bits)), (unsigned long long)emit_base);      // leaf = k-mer >> low_bits
		}
		if (failed) break;
		leaf = __shfl_sync(FULL, next_t, 0);
	}
	// ---- statistics of this warp
	if (failed) { if (lane == 0) atomicOr(a.flags, kMsdFlagFallback); return; }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		n_unique += __shfl_down_sync(FULL, n_unique, o);
		n_min += __shfl_down_sync(FULL, n_min, o);
		n_max += __shfl_down_sync(FULL, n_max, o);
	}
	if (lane == 0) {
		if (n_unique) atomicAdd(reinterpret_cast<unsigned long long*>(a.result), (unsigned long long)n_unique);
		if (n_min) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 1, (unsigned long long)n_min);
		if (n_max) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 2, (unsigned long long)n_max);
	}
}

// exclusive scan of the per-leaf record counts (single CTA; a warp owns 2048 consecutive leaves and scans them 32 at a time), total -> result[4]
__global__ void __launch_bounds__(1024) leaf_scan_kernel(const uint32_t* leaf_emit, uint32_t n_leaves, uint64_t* leaf_off, uint64_t* result, uint64_t out_capacity, uint32_t ob, const uint32_t* flags)
{
	__shared__ uint64_t s_w[32];
	if (*flags & kMsdFlagFallback) return;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t per = ((n_leaves + 1023) / 1024) * 32;          // leaves per warp (multiple of 32)
	const uint32_t w0 = warp * per;
	uint64_t sum = 0;
	for (uint32_t i = lane; i < per; i += 32) sum += (w0 + i < n_leaves) ? leaf_emit[w0 + i] : 0u;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
	if (lane == 0) s_w[warp] = sum;
	__syncthreads();
	uint64_t base = 0, tot = 0;
	for (uint32_t w = 0; w < 32; ++w) { if (w < warp) base += s_w[w]; tot += s_w[w]; }
	for (uint32_t i0 = 0; i0 < per; i0 += 32) {
		const uint32_t i = w0 + i0 + lane;
		const uint32_t v = i < n_leaves ? leaf_emit[i] : 0u;
		uint32_t inc = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= (uint32_t)o) inc += t;
		}
		if (i < n_leaves) leaf_off[i] = base + inc - v;
		base += __shfl_sync(0xffffffffu, inc, 31);
	}
	if (tid == 0) {
		result[4] = tot;
		if (tot * ob > out_capacity) result[5] = 1;
	}
}

// one warp per leaf: its padded temporary records -> packed records at their final place
__global__ void __launch_bounds__(256) leaf_gather_kernel(const uint8_t* tmp, const uint64_t* start, const uint32_t* leaf_emit, const uint64_t* leaf_off,
	uint32_t n_leaves, uint32_t ob, uint8_t* out, const uint64_t* result, const uint32_t* flags)
{
	if (*flags & kMsdFlagFallback) return;
	if (result[5]) return;                        // capacity error: nothing is written
	const uint32_t leaf = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
	if (leaf >= n_leaves) return;
	const uint32_t pad = ((ob + 7) >> 3) << 3;
	const uint32_t nbytes = leaf_emit[leaf] * ob;
	const uint8_t* src = tmp + start[leaf] * pad;
	uint8_t* dst = out + leaf_off[leaf] * ob;
	const uint32_t magic = 0xFFFFFFFFu / ob + 1;          // p / ob == umulhi(p, magic) for p < 2^16 ... checked: larger leaves take the division
	for (uint32_t p = lane; p < nbytes; p += 32) {
		const uint32_t r = nbytes < 65536u ? __umulhi(p, magic) : p / ob;
		dst[p] = src[(size_t)r * pad + (p - r * ob)];
	}
}

// when the hybrid path gave up after the leaves had already touched lut / result: start over for the fallback
__global__ void leaf_reset_kernel(uint64_t* lut, uint64_t lut_entries, uint64_t* result, const uint32_t* flags)
{
	if (!(*flags & kMsdFlagFallback)) return;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < lut_entries; i += (uint64_t)gridDim.x * blockDim.x) lut[i] = 0;
	if (blockIdx.x == 0 && threadIdx.x < 6) result[threadIdx.x] = 0;
}

}  // namespace kmcb
