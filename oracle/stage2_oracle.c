/* TEST INFRASTRUCTURE — see stage2_oracle.h.  Plain-C restatement of KMC 3.2.4 stage 2 for one bin.
 *
 * Every function cites the reference code it follows (paths relative to /root/reference).
 * The restatement always expands to plain canonical k-mers.  For k % 32 != 0 the reference goes
 * through (k,x)-mers instead (kmc.h:139-142, kb_sorter.h:371-638, 937-1122, kxmer_set.h); the emitted
 * bytes, LUT and statistics depend only on the multiset of canonical k-mers, which is
 * what tests/test_oracle_vs_reference.py verifies against the reference itself for k = 28, 55, 70 ...
 */
#include "stage2_oracle.h"
#include <stdlib.h>
#include <string.h>

uint32_t kmco_rec_words(uint32_t kmer_len) { return (kmer_len + 31) / 32; }

/* defs.h:121,154-159: BYTE_LOG and calc_counter_size */
static uint32_t byte_log(uint64_t x) { return x < (1u << 8) ? 1 : x < (1u << 16) ? 2 : x < (1u << 24) ? 3 : 4; }
uint32_t kmco_counter_bytes(const kmco_params* p)
{
	if (p->counter_max == 1) return 0;
	uint32_t a = byte_log(p->cutoff_max), b = byte_log(p->counter_max);
	return a < b ? a : b;
}
uint32_t kmco_out_rec_bytes(const kmco_params* p)
{
	return (p->kmer_len - p->lut_prefix_len) / 4 + kmco_counter_bytes(p);       /* kb_sorter.h:1132-1142 */
}

/* record walk: pos += 1 + (a + k + 3) / 4   (bkb_reader.cpp:48, kb_collector.cpp:34-90) */
uint64_t kmco_walk_bin(const uint8_t* data, uint64_t size, uint32_t k, uint64_t* n_rec)
{
	uint64_t pos = 0, n_sk = 0, n = 0;
	while (pos < size) {
		uint32_t a = data[pos];
		pos += 1 + (a + k + 3) / 4;
		n += a + 1;
		++n_sk;
	}
	if (pos != size) return (uint64_t)-1;
	if (n_rec) *n_rec = n;
	return n_sk;
}

/* ---- multi-word helpers: word 0 least significant (kmer.h:22-67) ---- */
static void shl2_insert(uint64_t* w, uint32_t W, uint32_t s)              /* kmer.h:223-239 SHL_insert_2bits */
{
	for (uint32_t i = W - 1; i > 0; --i) w[i] = (w[i] << 2) | (w[i - 1] >> 62);
	w[0] = (w[0] << 2) | s;
}
static void shr2_insert(uint64_t* w, uint32_t W, uint32_t s, uint32_t bitpos)   /* kmer.h:165-182 SHR_insert_2bits */
{
	for (uint32_t i = 0; i + 1 < W; ++i) w[i] = (w[i] >> 2) | (w[i + 1] << 62);
	w[W - 1] >>= 2;
	w[bitpos >> 6] |= (uint64_t)s << (bitpos & 63);
}
static void mask_k(uint64_t* w, uint32_t W, uint32_t k)                    /* kmer.h set_n_1 + mask */
{
	uint32_t bits = 2 * k;
	for (uint32_t i = 0; i < W; ++i) {
		if (bits >= 64 * (i + 1)) continue;
		if (bits <= 64 * i) w[i] = 0;
		else w[i] &= (~0ull) >> (64 - (bits - 64 * i));
	}
}
static int less_than(const uint64_t* a, const uint64_t* b, uint32_t W)     /* kmer.h:271-278 operator< : most significant word first */
{
	for (int i = (int)W - 1; i >= 0; --i) {
		if (a[i] < b[i]) return 1;
		if (a[i] > b[i]) return 0;
	}
	return 0;
}
static int equal(const uint64_t* a, const uint64_t* b, uint32_t W)
{
	for (uint32_t i = 0; i < W; ++i) if (a[i] != b[i]) return 0;
	return 1;
}
static uint32_t get_byte(const uint64_t* w, uint32_t j) { return (uint32_t)(w[j >> 3] >> ((j & 7) << 3)) & 0xFF; }   /* kmer.h:242-245 */

/* kb_sorter.h:251-298 (ExpandKmersAll) and :299-362 (ExpandKmersBoth).
 * Symbol i of a super-k-mer lives in byte i/4 after the length byte, bits 7-6 first (splitter/collector format). */
uint64_t kmco_expand(const kmco_params* p, const uint8_t* data, uint64_t size, uint64_t* recs)
{
	const uint32_t k = p->kmer_len, W = kmco_rec_words(k);
	uint64_t pos = 0, out = 0;
	uint64_t kmer[KMCO_MAX_WORDS], rev[KMCO_MAX_WORDS];
	while (pos < size) {
		uint32_t a = data[pos++];
		const uint8_t* sym = data + pos;
		uint32_t n = k + a;
		memset(kmer, 0, sizeof kmer);
		memset(rev, 0, sizeof rev);
		for (uint32_t i = 0; i < n; ++i) {
			uint32_t s = (sym[i >> 2] >> (6 - 2 * (i & 3))) & 3;
			shl2_insert(kmer, W, s);                           /* kb_sorter.h:353-354 */
			mask_k(kmer, W, k);
			shr2_insert(rev, W, 3 - s, 2 * (k - 1));           /* kb_sorter.h:355 */
			if (i + 1 >= k) {
				const uint64_t* c = (p->both_strands && !less_than(kmer, rev, W)) ? rev : kmer;   /* :340,356  kmer < rev ? kmer : rev */
				memcpy(recs + out * W, c, W * 8);
				++out;
			}
		}
		pos += (n + 3) / 4;
	}
	return out;
}

/* SortFunction contract (raduls.h:19-20): ascending on bytes key_bytes-1 .. 0 of the little-endian record image. */
void kmco_sort(uint64_t* recs, uint64_t* tmp, uint64_t n, uint32_t W, uint32_t key_bytes)
{
	uint64_t* src = recs; uint64_t* dst = tmp;
	uint64_t* cnt = (uint64_t*)malloc(256 * sizeof(uint64_t));
	for (uint32_t b = 0; b < key_bytes; ++b) {
		memset(cnt, 0, 256 * sizeof(uint64_t));
		for (uint64_t i = 0; i < n; ++i) cnt[get_byte(src + i * W, b)]++;
		if (n && cnt[get_byte(src, b)] == n) continue;         /* all equal in this byte: nothing to do */
		uint64_t s = 0;
		for (int d = 0; d < 256; ++d) { uint64_t c = cnt[d]; cnt[d] = s; s += c; }
		for (uint64_t i = 0; i < n; ++i) {
			uint64_t o = cnt[get_byte(src + i * W, b)]++;
			memcpy(dst + o * W, src + i * W, W * 8);
		}
		uint64_t* t = src; src = dst; dst = t;
	}
	if (src != recs) memcpy(recs, src, n * W * 8);
	free(cnt);
}

/* kmer.h:294-303 remove_suffix(2*(k-p)): the p leading symbols as an integer */
static uint64_t prefix_of(const uint64_t* w, uint32_t W, uint32_t nbits)
{
	uint32_t q = nbits >> 6, r = nbits & 63;
	if (q == W - 1 || r == 0) return w[q] >> r;
	return (w[q + 1] << (64 - r)) + (w[q] >> r);
}

/* kb_sorter.h:1128-1281 CompactKmers */
uint64_t kmco_compact(const kmco_params* p, const uint64_t* buf, uint64_t n_rec,
	uint8_t* out, uint64_t out_cap, uint64_t* lut, uint64_t stats[4])
{
	const uint32_t k = p->kmer_len, W = kmco_rec_words(k);
	const uint32_t kmer_symbols = k - p->lut_prefix_len;
	const uint32_t kmer_bytes = kmer_symbols / 4;
	const uint32_t counter_size = kmco_counter_bytes(p);
	const uint64_t lut_recs = 1ull << (2 * p->lut_prefix_len);
	uint64_t out_pos = 0, n_unique = 0, n_cutoff_min = 0, n_cutoff_max = 0, n_total = 0;
	if (lut) memset(lut, 0, lut_recs * 8);
	if (n_rec) {
		const uint64_t* act = buf;
		uint32_t count = 1;                                         /* uint32 like kb_sorter.h:1153 */
		n_total = n_rec;
		for (uint64_t i = 1; i <= n_rec; ++i) {
			if (i < n_rec && equal(act, buf + i * W, W)) { count++; continue; }
			/* run [act, i) finished (the tail after the loop, :1227-1267, is the same code) */
			if (count < p->cutoff_min) n_cutoff_min++;              /* :1174 */
			else if (count > p->cutoff_max) n_cutoff_max++;         /* :1181 */
			else {
				if (count > p->counter_max) count = p->counter_max; /* :1190 */
				if (out_pos + kmer_bytes + counter_size > out_cap) return (uint64_t)-1;
				for (int j = (int)kmer_bytes - 1; j >= 0; --j) out[out_pos++] = (uint8_t)get_byte(act, (uint32_t)j);   /* :1198 */
				for (uint32_t j = 0; j < counter_size; ++j) out[out_pos++] = (count >> (j * 8)) & 0xFF;                 /* :1200 */
				if (lut) lut[prefix_of(act, W, 2 * kmer_symbols)]++;                                                    /* :1203 */
			}
			n_unique++;
			if (i < n_rec) { act = buf + i * W; count = 1; }
		}
	}
	stats[0] = n_unique; stats[1] = n_cutoff_min; stats[2] = n_cutoff_max; stats[3] = n_total;
	return out_pos;
}

/* kb_sorter.h:210-237 */
uint64_t kmco_process_bin(const kmco_params* p, const uint8_t* data, uint64_t size, uint64_t n_rec,
	uint8_t* out, uint64_t out_cap, uint64_t* lut, uint64_t stats[4])
{
	const uint32_t W = kmco_rec_words(p->kmer_len);
	uint64_t* recs = (uint64_t*)malloc((n_rec + 1) * W * 8);
	uint64_t* tmp = (uint64_t*)malloc((n_rec + 1) * W * 8);
	uint64_t n = kmco_expand(p, data, size, recs);
	uint64_t r;
	if (n != n_rec) r = (uint64_t)-2;
	else {
		kmco_sort(recs, tmp, n, W, (p->kmer_len + 3) / 4);     /* rec_len, kb_sorter.h:769 */
		r = kmco_compact(p, recs, n, out, out_cap, lut, stats);
	}
	free(recs); free(tmp);
	return r;
}
