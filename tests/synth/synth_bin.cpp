// TEST / BENCH INFRASTRUCTURE - not part of the product (libkmc_b200.so does not contain it).
//
// Synthetic bins in stage 1's output format, i.e. what CKmerBinCollector::PutExtendedKmer writes (kmc_core/kb_collector.cpp:34-90):
// records [u8 a][ceil((k+a)/4) bytes, 2 bits per symbol, first symbol in bits 7-6], cut into expander packs of <= 64 KiB
// (one per collector flush, kb_collector.cpp:93-106).  Super-k-mers are substrings (random strand, `err_ppm` substitutions per
// million symbols) of a random genome of genome_len symbols, with `a` ~ geometric(mean mean_extra), capped at 255, until exactly
// n_rec k-mers exist: genome_len = n_rec / 30 gives the 30x duplicate-rich bins of the benchmark, genome_len >= n_rec all-distinct ones.
// Two-call protocol: with data == NULL only the sizes are returned.  Built by tests/synth/Makefile into tests/synth/libkmc_synth.so.
#include <stdint.h>
#include <cmath>
#include <vector>
#include <algorithm>

extern "C" {

static inline uint64_t splitmix64(uint64_t& x)
{
	uint64_t z = (x += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

int kmcsynth_bin(uint64_t seed, uint32_t k, uint64_t n_rec, uint64_t genome_len, double mean_extra, uint32_t err_ppm,
	uint8_t* data, uint64_t data_capacity, uint64_t* size, uint64_t* pack_bytes, uint64_t* pack_recs, uint32_t pack_capacity, uint32_t* n_packs, uint64_t* n_super_kmers)
{
	if (k < 1 || k > 256 || !size || !n_packs) return -1;
	if (genome_len < (uint64_t)k + 256) genome_len = (uint64_t)k + 256;
	uint64_t rs = seed * 0x2545F4914F6CDD1Dull + 0x1234567;
	std::vector<uint8_t> genome(genome_len);
	for (uint64_t i = 0; i < genome_len; i += 32) {
		uint64_t r = splitmix64(rs);
		for (uint64_t j = i; j < std::min(genome_len, i + 32); ++j) { genome[j] = r & 3; r >>= 2; }
	}
	const double pgeo = 1.0 / (mean_extra + 1.0);
	const double log1mp = std::log(1.0 - std::min(pgeo, 0.999999));
	// substitutions: the distance to the next one is geometric (one random number per error, not per symbol)
	const double perr = (double)err_ppm * 1e-6;
	const double log1me = perr > 0 && perr < 1 ? std::log(1.0 - perr) : 0.0;
	auto next_gap = [&]() -> uint64_t {
		if (perr <= 0) return ~0ull;
		if (perr >= 1) return 0;
		const double u = (double)((splitmix64(rs) >> 11) + 1) * (1.0 / 9007199254740992.0);
		return (uint64_t)(std::log(u) / log1me);
	};
	uint64_t to_err = next_gap();
	uint64_t pos_out = 0, made = 0, n_sk = 0;
	uint32_t np = 0;
	uint64_t pack_fill = 0, pack_kmers = 0;          // pack_recs[i] = k-mers of pack i (CExpanderPackDesc's second field is an upper bound of its (k+x)-mers)
	uint8_t symbuf[256 + 256 + 8];
	while (made < n_rec) {
		double u = (double)(splitmix64(rs) >> 11) * (1.0 / 9007199254740992.0);
		uint64_t a = pgeo >= 0.999999 ? 0 : (uint64_t)(std::log(1.0 - u) / log1mp);
		if (a > 255) a = 255;
		if (a + 1 > n_rec - made) a = n_rec - made - 1;
		const uint32_t n = k + (uint32_t)a;
		const uint64_t p = splitmix64(rs) % (genome_len - n + 1);
		const bool rc = splitmix64(rs) & 1;
		for (uint32_t i = 0; i < n; ++i) {
			uint8_t s = rc ? (uint8_t)(3 - genome[p + n - 1 - i]) : genome[p + i];
			if (to_err-- == 0) { s = (uint8_t)((s + 1 + splitmix64(rs) % 3) & 3); to_err = next_gap(); }
			symbuf[i] = s;
		}
		const uint32_t bytes = 1 + (n + 3) / 4;
		if (pack_fill + bytes > (1u << 16)) {               // collector flush (kb_collector.cpp:44-55)
			if (pack_bytes && np < pack_capacity) pack_bytes[np] = pack_fill;
			if (pack_recs && np < pack_capacity) pack_recs[np] = pack_kmers;
			++np;
			pack_fill = 0; pack_kmers = 0;
		}
		if (data) {
			if (pos_out + bytes > data_capacity) return -5;
			data[pos_out] = (uint8_t)a;
			for (uint32_t i = 0; i < (n + 3) / 4; ++i) {
				uint8_t b = 0;
				for (uint32_t j = 0; j < 4; ++j) { const uint32_t q = 4 * i + j; b = (uint8_t)((b << 2) | (q < n ? symbuf[q] : 0)); }
				data[pos_out + 1 + i] = b;
			}
		}
		pos_out += bytes; pack_fill += bytes; pack_kmers += a + 1; made += a + 1; ++n_sk;
	}
	if (pack_fill) { if (pack_bytes && np < pack_capacity) pack_bytes[np] = pack_fill; if (pack_recs && np < pack_capacity) pack_recs[np] = pack_kmers; ++np; }
	*size = pos_out; *n_packs = np;
	if (n_super_kmers) *n_super_kmers = n_sk;
	if (pack_bytes && np > pack_capacity) return -5;
	return 0;
}

}  // extern "C"
