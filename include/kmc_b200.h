/* kmc_b200 — C ABI of the B200 (sm_100a) implementation of KMC's per-bin stage 2.
 *
 * This is the drop-in boundary: a KMC build binds these entry points where its own CPU code does the
 * per-bin work today.  File:line references are to refresh-bio/KMC 3.2.4.
 *
 *   seam #2 (the drop-in)  kmcb200_process_bin / kmcb200_submit_bin + kmcb200_wait_bin
 *        replaces the body of CKmerBinSorter<SIZE>::ProcessBins  (kmc_core/kb_sorter.h:210-237):
 *        Expand (:728-752) -> Sort (:757-780) -> Compact (:1287-1293) for one bin, i.e. everything
 *        between sorters_manager->GetNext()/bd->read() and kq->push().
 *   seam #1 (sort only)    kmcb200_sort_records
 *        replaces SortFunction<CKmer<SIZE>> (kmc_core/raduls.h:19-20), i.e. RadulsSort::RadixSortMSD_*
 *        (raduls_impl.h:769-776) / RadixSort::RadixSortMSD (radix.h:845-855), as called at kb_sorter.h:775.
 *   device-level twins     kmcb200_dev_*  — same operations on buffers that already live in HBM
 *        (used by bench.py for the kernel-only numbers and by hosts that keep bins resident).
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative kmcb200_status;
 * kmcb200_last_error() gives the message.  A context is bound to one GPU and may be used by one host
 * thread at a time (KMC runs one sorter thread per context).  There is NO CPU fallback: without a usable
 * sm_100 device kmcb200_create fails with KMCB200_ERR_NO_DEVICE.
 *
 * Environment knobs read by kmcb200_create (development / tests; the defaults are the measured best):
 *   KMCB200_SORT=lsd               plain 8-bit LSD passes instead of the hybrid MSD sort
 *   KMCB200_LEAF=sort              sort the leaves on chip + count_emit instead of counting them in hash tables
 *   KMCB200_LEAF_SLOT_BITS=8|9|10  slots of a warp's leaf table (default 10)
 *   KMCB200_LEAF_KERNEL=warp       round 1's leaf kernel (ordered groups, leaf_warp.cuh) instead of leaf_hash_kernel; KMCB200_LEAF_WIDE=warp: for records of > 1 word only
 *   KMCB200_LEAF_FILL_PCT=n        leaf_hash_kernel plans a table round for this load (default 62); KMCB200_LEAF_RATIO0=n: first guess of distinct k-mers per record x 256 (default 90)
 *   KMCB200_LEAF_ROUND_PCT=n       leaf_warp_kernel: records per table round in percent of the slots (default 100)
 *   KMCB200_L2_BITS=1..10          bits of the second partition level (default: from the bin size: 8, from ~10^8 one-word k-mers on 9; wider records up to 10)
 *   KMCB200_LEAF_TARGET=n, KMCB200_LEAF_MAX_B2=8..10   mean leaf size / most bits the default rule aims at for one-word records (1024, 9)
 *   KMCB200_MAX_BLOCK_RECORDS=n    a bin with more k-mers is counted key block by key block (default: what 60 % of the free HBM holds, < 2^32)
 *   KMCB200_KEY_BLOCKS=filter     key blocks of an oversized bin re-expand it with a filter (default: one scattering expansion when the records fit in HBM once)
 *   KMCB200_KEY_BLOCK_RECORDS=n    preferred size of a key block in the scattering flow (default 2^28)
 *   KMCB200_EXPAND=fused           single-pass expansion (expand_fused.cuh) instead of the index-based kernels (measured slower; option)
 *   KMCB200_OVERLAP_WALK=0         index kernels of a submitted bin on the compute stream instead of its copy stream
 *   KMCB200_MAX_CHUNK_BYTES=n      ... and expanded in chunks of at most n bytes (default 2^30)
 */
#ifndef KMC_B200_H
#define KMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMCB200_VERSION 1
#define KMCB200_MAX_KMER_LEN 128          /* records of up to 4 x 64 bit */
#define KMCB200_MAX_SLOTS 4

typedef enum {
	KMCB200_OK = 0,
	KMCB200_ERR_INVALID = -1,             /* bad argument */
	KMCB200_ERR_NO_DEVICE = -2,           /* no CUDA device / not sm_100 */
	KMCB200_ERR_CUDA = -3,                /* CUDA runtime error (message in last_error) */
	KMCB200_ERR_BIN_FORMAT = -4,          /* packs do not end on record boundaries / n_rec mismatch */
	KMCB200_ERR_CAPACITY = -5,            /* out_capacity too small */
	KMCB200_ERR_BUSY = -6                 /* slot already holds a submitted bin */
} kmcb200_status;

typedef struct kmcb200_ctx kmcb200_ctx;

/* Per-run parameters: the fields CKmerBinSorter's constructor takes from CKMCParams (kb_sorter.h:165-200). */
typedef struct {
	uint32_t kmer_len;                    /* Params.kmer_len, 1..KMCB200_MAX_KMER_LEN */
	uint32_t both_strands;                /* Params.both_strands: 1 = canonical k-mers */
	uint32_t cutoff_min;                  /* Params.cutoff_min */
	uint32_t cutoff_max;                  /* (uint32)Params.cutoff_max   (kb_sorter.h:186) */
	uint32_t counter_max;                 /* (uint32)Params.counter_max  (kb_sorter.h:187) */
	uint32_t lut_prefix_len;              /* Params.lut_prefix_len, >= 1, (kmer_len - lut_prefix_len) % 4 == 0 */
	int32_t device;                       /* CUDA ordinal */
	uint32_t n_slots;                     /* bins in flight per context (1..KMCB200_MAX_SLOTS); 2 overlaps copies with kernels */
} kmcb200_params;

int kmcb200_create(const kmcb200_params* params, kmcb200_ctx** out_ctx);
void kmcb200_destroy(kmcb200_ctx* ctx);
/* message of the last failure on this context (or of the last failed kmcb200_create when ctx == NULL) */
const char* kmcb200_last_error(const kmcb200_ctx* ctx);

/* bytes of one emitted database record: (k-p)/4 suffix bytes + counter bytes (kb_sorter.h:1132-1142, defs.h:154-159) */
uint32_t kmcb200_out_rec_bytes(const kmcb200_ctx* ctx);
/* size in bytes of the out buffer the reference reserves for a bin of n_rec k-mers (kb_reader.h:141-150) */
uint64_t kmcb200_out_capacity(const kmcb200_ctx* ctx, uint64_t n_rec);
/* entries of the per-bin LUT: 4^lut_prefix_len */
uint64_t kmcb200_lut_entries(const kmcb200_ctx* ctx);

/* Pinned host memory for bin / result buffers (cudaHostAlloc).  A host may instead cudaHostRegister its own arena. */
int kmcb200_host_alloc(kmcb200_ctx* ctx, uint64_t bytes, void** out_ptr);
int kmcb200_host_free(kmcb200_ctx* ctx, void* ptr);

/* ---- seam #2: one bin, host buffers --------------------------------------------------------------
 * Inputs  = what CKmerBinSorter::ProcessBins gets from sorters_manager->GetNext / bd->read / epd->pop:
 *   superkmers/size : the bin byte stream (CMemoryBins::mba_input_file)        kb_sorter.h:216
 *   n_rec           : number of k-mers in the bin (CBinDesc)                   kb_sorter.h:219
 *   n_plus_x_recs   : (k+x)-mer estimate; accepted for signature parity, unused (we never build (k,x)-mers)
 *   pack_bytes[n_packs] : byte length of every expander pack (CExpanderPackDesc, first of each pair,
 *                     queues.h:376-396); packs start on record boundaries.  pack_recs may be NULL (unused).
 * Outputs = what is handed to kq->push (kb_sorter.h:1273, queues.h:826):
 *   out_suffix[0, *out_bytes) : the emitted records (one data pack (0, out_pos))
 *   lut[4^p]                  : raw per-prefix counts (the completer does the prefix sum)
 *   stats[4]                  : n_unique, n_cutoff_min, n_cutoff_max, n_total
 * An empty bin (size == 0) is legal and yields zero output and a zero LUT (kb_reader.h:198-205).
 */
int kmcb200_process_bin(kmcb200_ctx* ctx, int32_t bin_id,
	const uint8_t* superkmers, uint64_t size, uint64_t n_rec, uint64_t n_plus_x_recs,
	const uint64_t* pack_bytes, const uint64_t* pack_recs, uint32_t n_packs,
	uint8_t* out_suffix, uint64_t out_capacity, uint64_t* out_bytes,
	uint64_t* lut, uint64_t stats[4]);

/* Asynchronous form: submit returns once the work is queued on the slot's stream; wait blocks until the
 * bin's outputs are in the host buffers given to submit.  Host buffers should be pinned for real overlap. */
int kmcb200_submit_bin(kmcb200_ctx* ctx, uint32_t slot, int32_t bin_id,
	const uint8_t* superkmers, uint64_t size, uint64_t n_rec, uint64_t n_plus_x_recs,
	const uint64_t* pack_bytes, const uint64_t* pack_recs, uint32_t n_packs,
	uint8_t* out_suffix, uint64_t out_capacity, uint64_t* lut);
int kmcb200_wait_bin(kmcb200_ctx* ctx, uint32_t slot, uint64_t* out_bytes, uint64_t stats[4]);

/* SURVEY 8f N4 - a device-friendly stage-1 output: kmcb200_submit_bin for a stage 1 that also hands over, per bin, the length byte `a` of
 * every record as a separate array (extras[n_super_kmers], stream order: the value CKmerBinCollector::PutExtendedKmer stores in front of the
 * record, kb_collector.cpp:60-66) and the number of records of every expander pack (pack_superkmers[n_packs]).  The record index is then two
 * parallel prefix sums per pack instead of the serial record walk; it is verified against the stream, a wrong array is KMCB200_ERR_BIN_FORMAT.
 * INTEGRATION.md shows the few lines a KMC collector needs for it. */
int kmcb200_submit_bin_indexed(kmcb200_ctx* ctx, uint32_t slot, int32_t bin_id,
	const uint8_t* superkmers, uint64_t size, uint64_t n_rec, const uint64_t* pack_bytes, uint32_t n_packs,
	const uint8_t* extras, uint64_t n_super_kmers, const uint32_t* pack_superkmers,
	uint8_t* out_suffix, uint64_t out_capacity, uint64_t* lut);

/* One bin over several GPUs (the reference's analogue is RADULS' team sort of a big bucket, raduls_impl.h:672-745): for a bin
 * that is too large for a fair share of one GPU's time.  ctxs[0..n_ctx) are contexts with identical parameters on different (or, for
 * tests, the same) devices, none with a bin in flight; one host thread per GPU is started inside the call.  Every GPU gets a contiguous
 * share of the bin's packs from the host and counts its top 12 bits; the host cuts the key space into one range per GPU and the ranges into
 * key blocks; every GPU expands its share once, scattering the k-mers by key block; the records are exchanged with peer copies (NVLink:
 * an all-to-all of 8 B x n_rec x (N-1)/N); every GPU sorts and counts its own blocks.  Outputs are concatenated in key order:
 * byte-identical to kmcb200_process_bin on one GPU. */
int kmcb200_process_bin_multi(kmcb200_ctx* const* ctxs, uint32_t n_ctx, int32_t bin_id,
	const uint8_t* superkmers, uint64_t size, uint64_t n_rec, const uint64_t* pack_bytes, uint32_t n_packs,
	uint8_t* out_suffix, uint64_t out_capacity, uint64_t* out_bytes, uint64_t* lut, uint64_t stats[4]);

/* ---- database assembly without the reference's completer loop (kmc_core/kb_completer.cpp:59-326) ------------------------
 * kmcb200_wait_bin_scanned = kmcb200_wait_bin, but the LUT arrives as the completer writes it to .kmc_pre: the exclusive prefix sum of the
 * bin's raw counts offset by lut_base (records preceding the bin in the file, kb_completer.cpp:191-201), computed on the GPU before the copy.
 * The writer owns a PINNED staging ring: kmcb200_db_reserve gives the region the next bin's records are copied into by the GPU (pass it as
 * out_suffix of kmcb200_submit_bin), kmcb200_db_commit_bin queues it for the writer thread (fwrite to .kmc_suf / .kmc_pre in commit order,
 * overlapping the GPU), kmcb200_db_close writes the footer of ProcessBinsSecondStage (:284-320).  For the same bins in the same order the
 * two files are byte-identical to the reference's. */
typedef struct kmcb200_db_writer kmcb200_db_writer;
typedef struct {
	uint32_t kmer_len, counter_size /* bytes */, lut_prefix_len, signature_len, cutoff_min, cutoff_max, both_strands;
} kmcb200_db_params;
int kmcb200_wait_bin_scanned(kmcb200_ctx* ctx, uint32_t slot, uint64_t lut_base, uint64_t* out_bytes, uint64_t stats[4]);
int kmcb200_db_open(const kmcb200_db_params* params, const char* path_prefix, uint64_t staging_bytes, kmcb200_db_writer** out_writer);
const char* kmcb200_db_last_error(const kmcb200_db_writer* w);
uint64_t kmcb200_db_records(const kmcb200_db_writer* w);           /* records committed so far = lut_base of the next bin */
int kmcb200_db_reserve(kmcb200_db_writer* w, uint64_t bytes, uint8_t** out_ptr);
int kmcb200_db_commit_bin(kmcb200_db_writer* w, uint64_t payload_bytes, const uint64_t* lut, int raw_lut, const uint64_t stats[4],
	const uint32_t* signatures, uint32_t n_signatures);
int kmcb200_db_close(kmcb200_db_writer* w, uint64_t totals[4]);

/* ---- seam #1: sort host records ------------------------------------------------------------------
 * Contract of SortFunction (raduls.h:19-20, kb_sorter.h:775-779): n records of rec_bytes (multiple of 8,
 * CKmer<SIZE> images) sorted ascending on bytes key_bytes-1..0; the result is left in `tmp` when key_bytes
 * is odd and in `recs` when it is even.  Returns 1 / 0 for tmp / recs, negative on error. */
int kmcb200_sort_records(kmcb200_ctx* ctx, void* recs, void* tmp, uint64_t n, uint32_t rec_bytes, uint32_t key_bytes);

/* ---- device-level entry points (pointers are DEVICE pointers, `stream` is a cudaStream_t or NULL) ---
 * All work is enqueued on `stream` (NULL = the context's compute stream) and is asynchronous with respect to the host. */

/* Expand + sort + count one bin that already lives in HBM.  d_superkmers must be 8-byte aligned and READABLE FOR 32 BYTES PAST
 * `size` (the walk and the staging use 16-byte vector loads on the absolute 16-byte grid); likewise the record buffers given to
 * kmcb200_dev_sort / kmcb200_dev_count must be readable for one record past n (the TMA tile loads round an odd count up to an
 * even one).  The values read there are never used.  pack_bytes is a HOST array.  d_result receives 8 x uint64:
 * [0..3] stats, [4] emitted records, [5] capacity error flag, [6] bin-format error bits,
 * [7] 1 when the hybrid MSD / leaf-count path gave up (skew) and the LSD fallback produced the (identical) result. */
int kmcb200_dev_process_bin(kmcb200_ctx* ctx, uint32_t slot,
	const uint8_t* d_superkmers, uint64_t size, uint64_t n_rec,
	const uint64_t* pack_bytes, uint32_t n_packs,
	uint8_t* d_out, uint64_t out_capacity, uint64_t* d_lut, uint64_t* d_result, void* stream);

/* Individual stages, for per-kernel measurement.  d_recs/d_tmp hold n records of 8*ceil(k/32) bytes.
 * kmcb200_dev_expand also leaves the first partition level's work items and digit counts in the slot's workspace, which
 * kmcb200_dev_sort(..., hist_ready=1) consumes; with hist_ready=0 the sort counts the first digit itself.
 * kmcb200_dev_sort returns 1 when the sorted records are in d_tmp, 0 when they are in d_recs. */
int kmcb200_dev_expand(kmcb200_ctx* ctx, uint32_t slot, const uint8_t* d_superkmers, uint64_t size, uint64_t n_rec,
	const uint64_t* pack_bytes, uint32_t n_packs, void* d_recs, uint64_t* d_result, void* stream);
int kmcb200_dev_sort(kmcb200_ctx* ctx, uint32_t slot, void* d_recs, void* d_tmp, uint64_t n, uint32_t key_bytes,
	int hist_ready, void* stream);
int kmcb200_dev_count(kmcb200_ctx* ctx, uint32_t slot, const void* d_sorted, uint64_t n,
	uint8_t* d_out, uint64_t out_capacity, uint64_t* d_lut, uint64_t* d_result, void* stream);

/* Number of kernels this library has launched on the context so far (bench.py reports the delta). */
uint64_t kmcb200_kernel_launches(const kmcb200_ctx* ctx);
/* Duration in ms of the last-run stages of a slot, measured with CUDA events on the launching stream:
 * ms[0] index+expand, ms[1] sort (all of it), ms[2] count/emit, ms[3..3+n) the n timed intervals of the sort (hybrid MSD:
 * level-1 partition, level-2 count, level-2 partition, leaves, LSD fallback; or the plain LSD passes).
 * Blocks until the slot's work has finished.  Returns n, negative on error. */
int kmcb200_stage_times(kmcb200_ctx* ctx, uint32_t slot, float* ms, uint32_t capacity);
/* Comma-separated names of the timed sort intervals ms[3..] of kmcb200_stage_times ("msd_partition_L1,msd_count_L2,...").
 * Returns their number. */
int kmcb200_stage_names(kmcb200_ctx* ctx, uint32_t slot, char* buf, uint32_t capacity);

#ifdef __cplusplus
}
#endif
#endif
