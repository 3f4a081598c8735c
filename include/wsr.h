/* wsr.h — C ABI of the B200-native query engine for WiSER/Vacuum's hot path:
 * conjunctive posting-list intersection + fused BM25 scoring + top-k over an HBM-resident
 * inverted index (libwsr.so, hand-written sm_100a CUDA kernels; no CPU fallback).
 *
 * The reference has no FFI for this path; its seam is the C++ abstract class
 * SearchEngineServiceNew (reference src/qq_mem/src/engine_services.h:14-27) created by
 * CreateSearchEngine (engine_factory.h:33-50). Each entry point below names the reference
 * member it stands behind; the C++ adapter a maintainer adds on the reference side
 * (class GpuVacuumEngine : SearchEngineServiceNew) is shown in INTEGRATION.md and shipped as
 * wiser_b200/csrc/gpu_vacuum_engine.h.
 *
 * Conventions: plain pointers and sizes only; the caller owns every in/out buffer, the
 * library owns device memory and the handles. All functions return 0 on success or a
 * negative wsr_status; wsr_last_error() gives the message for the calling thread.
 * Thread-safety: a wsr_index is read-only after open; wsr_search / wsr_search_batch may be
 * called concurrently from many threads (the reference's server calls Search from N threads
 * on one shared engine with no lock, grpc_server_impl.h:309-328).
 */
#ifndef WSR_H
#define WSR_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define WSR_MAX_TERMS 8      /* QueryProcessor's phrase capacity, query_processing.h:695 */
#define WSR_TERM_ABSENT 0xFFFFFFFFu
#define WSR_QUERY_PHRASE 1u
#define WSR_OPEN_POSITIONS 1u   /* also load the position column (4 B per token in HBM) */

typedef enum {
  WSR_OK = 0,
  WSR_ERR_ARG = -1,        /* bad argument */
  WSR_ERR_IO = -2,         /* cannot read / parse the index directory */
  WSR_ERR_CUDA = -3,       /* CUDA runtime error or no device */
  WSR_ERR_UNSUPPORTED = -4 /* e.g. more than WSR_MAX_TERMS terms */
} wsr_status;

typedef struct wsr_index wsr_index;  /* one HBM-resident index (or document shard of one) */
typedef struct wsr_batch wsr_batch;  /* device-resident query batch + result buffers */

/* One result entry = SearchResultEntry{doc_id:int, doc_score:double} (types.h:259-263),
 * snippet omitted (return_snippets=false on this path). 16 bytes. */
typedef struct {
  int32_t doc_id;
  int32_t reserved;
  double score;
} wsr_hit;

/* One query = SearchQuery{terms, n_results} (types.h:205-218) after term lookup.
 * term_ids are in QUERY ORDER (BM25 partial scores are summed in that order, scoring.h:124-145).
 * A term id of WSR_TERM_ABSENT, or k == 0, makes the result empty with n_doc_freqs == 0
 * (vacuum_engine.h:206-215). */
typedef struct {
  uint32_t term_ids[WSR_MAX_TERMS];
  uint32_t n_terms;
  uint32_t k;      /* n_results */
  uint32_t flags;  /* WSR_QUERY_PHRASE: SearchQuery::is_phrase — with >= 2 terms a document is
                    * ranked only if the terms occur at consecutive positions, in query order
                    * (query_processing.h:886-895); needs WSR_OPEN_POSITIONS */
} wsr_query;

/* Index statistics. */
typedef struct {
  int64_t n_docs;            /* DocLengthCharStore::Size(): global, all shards */
  double avg_doc_len;        /* stored running mean from my.doc_length (never recomputed) */
  int64_t n_terms;           /* VacuumEngine::TermCount() */
  int64_t n_postings;        /* postings resident on THIS shard */
  int64_t n_postings_global; /* postings of the whole index */
  int64_t n_blocks;          /* 128-posting blocks on this shard */
  int64_t hbm_bytes;         /* device memory held by the index */
  int64_t payload_bytes;     /* packed doc-id + tf bitstreams */
  int32_t shard, n_shards;
  int32_t doc_lo, doc_hi;    /* this shard holds docs in [doc_lo, doc_hi) */
  int32_t device;
} wsr_index_info;

const char *wsr_last_error(void);
int wsr_device_count(void);
/* Page-locked host memory: result/query buffers allocated here are DMA'd directly by
 * wsr_search_batch / wsr_batch_fetch; other buffers go through an internal pinned stage. */
void *wsr_host_alloc(size_t bytes);
void wsr_host_free(void *p);

/* ---- Load(): VacuumEngine::Load (vacuum_engine.h:144-180) --------------------------------
 * Reads <dir>/my.tip, my.vacuum, my.doc_length and re-lays the posting lists into HBM on
 * `device` as fixed 128-posting packed delta blocks with per-block skip metadata.
 * Document partitioning: shard s of n_shards keeps the postings whose doc id falls in
 * [s*N/n, (s+1)*N/n); idf and doc_freqs keep using GLOBAL N and df. Use (0,1) for no
 * sharding. */
wsr_index *wsr_index_open(const char *vacuum_dir, int device, int shard, int n_shards,
                          int loader_threads, char *err, size_t errlen);
/* Same with flags: WSR_OPEN_POSITIONS loads the reference's position column (skip-list
 * addressed "cozy box" packs, flash_iterators.h:558-661) so that phrase queries can be served. */
wsr_index *wsr_index_open_ex(const char *vacuum_dir, int device, int shard, int n_shards,
                             int loader_threads, unsigned flags, char *err, size_t errlen);
void wsr_index_close(wsr_index *idx);
int wsr_index_get_info(const wsr_index *idx, wsr_index_info *info);

/* ---- document-partitioned deployment -----------------------------------------
 * Every GPU opens ITS OWN partition directory (a standalone vacuum index whose doc ids start at
 * 0) and is then told the collection-wide statistics, because the reference scores with the
 * whole index's N, average length and document frequencies:
 *   doc_base       global id of this partition's doc 0 (added to every emitted doc id)
 *   n_docs_global  N used by idf (calc_es_idf, scoring.h:21-25)
 *   avg_len_global average document length used by the BM25 length-norm cache (scoring.h:85-90)
 *   df_global      per LOCAL term id, the term's document frequency over all partitions
 * Recomputes idf, the 256-entry cache and every block's block-max bound on the device. */
int wsr_index_set_global_stats(wsr_index *idx, int64_t doc_base, int64_t n_docs_global,
                               double avg_len_global, const uint32_t *df_global);
/* Shard-local statistics for that exchange: df_local[i] = postings of term i on this shard.
 * For synthetic corpora whose terms are named t<rank>, ranks[i] = that rank (else the call
 * fails with WSR_ERR_ARG); either pointer may be NULL. */
int wsr_index_local_stats(const wsr_index *idx, uint32_t *df_local, uint32_t *ranks);

/* ---- term dictionary: TermTrieIndex::Find (term_index.h:136-144) -------------------------
 * Returns 0 and fills term_id / df (GLOBAL document frequency = posting-list size,
 * VacuumEngine::PostinglistSizes) or 1 if the term is absent. */
int wsr_term_lookup(const wsr_index *idx, const char *term, size_t len, uint32_t *term_id,
                    uint32_t *df);
/* i-th term in my.tip order; returns its length (copies at most cap bytes), <0 on error. */
int wsr_term_at(const wsr_index *idx, uint32_t term_id, char *buf, size_t cap, uint32_t *df);

/* Query-log text -> wsr_query records (QueryProducerByLog's loader, query_pool.h:314-335):
 * one query per line, terms separated by single spaces, a line wrapped in double quotes is a
 * phrase (flags bit 0). Terms are looked up in the dictionary; absent terms get
 * WSR_TERM_ABSENT. Every query gets n_results = k. */
int wsr_parse_query_log(const wsr_index *idx, const char *text, size_t len, int k,
                        wsr_query *out, int cap, int *n_out);

/* ---- decode: VacuumPostingListIterator walk (flash_iterators.h:985-1016) -----------------
 * Decodes this shard's part of a posting list on the GPU into HOST buffers (doc ids and
 * term frequencies in list order). *n receives the number of postings on this shard; at
 * most cap are written. */
int wsr_decode_list(wsr_index *idx, uint32_t term_id, uint32_t *docs, uint32_t *tfs,
                    size_t cap, size_t *n);
/* Bench/ncu entry: decodes EVERY block of the shard into a device scratch checksum without
 * copying results back; *checksum = sum of all doc ids + tfs (mod 2^64). */
int wsr_decode_all(wsr_index *idx, uint64_t *checksum, float *kernel_ms);

/* ---- Search(): VacuumEngine::Search (vacuum_engine.h:201-258) ----------------------------
 * String-term single query, blocking; flags = WSR_QUERY_PHRASE for SearchQuery::is_phrase.
 * hits must hold k entries, doc_freqs n_terms entries.
 * *n_doc_freqs is 0 when the reference returns early with an empty result, else n_terms. */
int wsr_search(wsr_index *idx, const char *const *terms, const size_t *term_lens, int n_terms,
               int k, unsigned flags, wsr_hit *hits, int *n_hits, uint32_t *doc_freqs,
               int *n_doc_freqs);

/* Batched Search over HOST buffers: the query batch is copied to the device, processed by
 * the batch scheduler in one pass, and the results copied back (all inside the call). Batches
 * of >= 8192 queries with k_stride <= 32 are planned on the GPU, smaller ones by host threads.
 * hits: n * k_stride entries (query i at hits[i*k_stride]); n_hits: n entries;
 * doc_freqs (may be NULL): n * WSR_MAX_TERMS entries, n_doc_freqs (may be NULL): n entries. */
int wsr_search_batch(wsr_index *idx, const wsr_query *queries, int n, int k_stride,
                     wsr_hit *hits, int32_t *n_hits, uint32_t *doc_freqs,
                     int32_t *n_doc_freqs);

/* Replays a whole query log given as TEXT — the replay driver's inner loop: QueryProducerByLog
 * (query_pool.h:251-352: one query per line, trimmed, a "quoted" line is a phrase, terms split on
 * ' ') + TermTrieIndex::Find (term_index.h:136-144) + VacuumEngine::Search per line, in one
 * blocking call. For k <= 32 only the text crosses PCIe: it is parsed, looked up in the HBM copy
 * of the term dictionary and planned by kernels (csrc/frontend.cu); larger k (or
 * WSR_HOST_FRONTEND=1) parse and plan on host threads, in up to 4 chunks overlapped with the GPU.
 * hits: cap_q * k entries (query i at hits[i*k], entries past n_hits[i] unspecified); n_hits:
 * cap_q entries; *n_queries receives the number of log lines. Pinned buffers (wsr_host_alloc) are
 * written by the kernels themselves, a row when its query finishes (no result copy behind the
 * kernels); other buffers are filled through a pinned stage (sparse results packed on the GPU
 * first). A line with more than WSR_MAX_TERMS terms fails the call. */
int wsr_search_log(wsr_index *idx, const char *text, size_t len, int k, wsr_hit *hits,
                   int32_t *n_hits, int cap_q, int *n_queries);
/* Same with SearchResult::doc_freqs (vacuum_engine.h:217-219): doc_freqs[i*WSR_MAX_TERMS + t] =
 * collection-wide df of term t of line i, n_doc_freqs[i] = its number of terms, 0 where the
 * reference returns early (n_results == 0, no terms, a term missing from the dictionary). Both
 * arrays hold cap_q rows; pass NULL for both to skip them. */
int wsr_search_log_ex(wsr_index *idx, const char *text, size_t len, int k, wsr_hit *hits,
                      int32_t *n_hits, uint32_t *doc_freqs, int32_t *n_doc_freqs, int cap_q,
                      int *n_queries);

/* ---- device-resident batches (replay driver, multi-GPU merge, benchmarking) --------------
 * wsr_batch_create uploads and plans a batch once; wsr_batch_run launches the kernels on the
 * batch's stream (no host<->device copies); results stay in device memory until fetched. */
wsr_batch *wsr_batch_create(wsr_index *idx, const wsr_query *queries, int n, int k_stride);
/* Re-plans and re-uploads an existing batch object with a new set of queries (buffers reused). */
int wsr_batch_reset(wsr_batch *b, const wsr_query *queries, int n, int k_stride);
/* Same from query-log text (the wsr_search_log front end: for k <= 32 the text is copied to the
 * GPU and parsed, looked up and planned there). *n_queries receives the number of log lines;
 * the batch's k_stride becomes k. */
int wsr_batch_reset_log(wsr_batch *b, const char *text, size_t len, int k, int *n_queries);
void wsr_batch_destroy(wsr_batch *b);
int wsr_batch_run(wsr_batch *b);                       /* asynchronous on the batch stream */
int wsr_batch_sync(wsr_batch *b);
int wsr_batch_fetch(wsr_batch *b, wsr_hit *hits, int32_t *n_hits); /* D2H + sync */
/* Device pointers of the result arrays (n*k_stride wsr_hit, n int32) and the CUDA stream
 * (cudaStream_t) — for NCCL allgather of per-shard top-k by the caller. */
int wsr_batch_device_results(wsr_batch *b, void **d_hits, void **d_n_hits, void **stream);
/* Timed run: launches the batch `iters` times back to back and reports the average device
 * time per iteration measured with CUDA events on the batch stream. */
int wsr_batch_time(wsr_batch *b, int iters, float *ms_per_iter);

/* One pass with CUDA events around every launch group; ms[0..3] = search kernels of the
 * single-term / two-term / 3+-term / collect classes, ms[4] = merge + collect epilogue,
 * ms[5] = whole pass. */
int wsr_batch_profile(wsr_batch *b, float ms[6]);

/* One pass of the batch through the kernel instantiations that COUNT their work (decoded
 * postings, touched algorithmic bytes, matches); same results as wsr_batch_run. The bookkeeping
 * costs ~7 % of the two-term kernel, so ordinary runs are compiled without it. */
int wsr_batch_count_work(wsr_batch *b);

/* Counters of the LAST run of this batch (read back from the device). decoded_postings,
 * touched_bytes and matches are 0 unless that run was wsr_batch_count_work. */
typedef struct {
  uint64_t listed_postings;   /* sum over valid queries of sum_i df_i (shard-local lists) */
  uint64_t decoded_postings;  /* postings of the blocks decoded IN FULL: driver blocks, the partner
                               * blocks of the merge path, the blocks a single-term query scores */
  uint64_t touched_bytes;     /* B_touched: per block decoded in full its doc-id pack
                               * (+ its tf pack where tfs are read with it) + 16 B metadata, per
                               * probed partner block (once per work unit) its doc-id pack + 16 B,
                               * + 1 norm byte per intersection hit, + 4 B per block-max entry scanned
                               * (+ 4 B per position read by phrase queries) */
  uint64_t listed_bytes;      /* algorithmic bytes of every block of every listed list */
  uint64_t matches;           /* intersection hits scored */
  uint64_t work_units;        /* warp work units scheduled */
  uint32_t kernel_launches;   /* launches per run */
  uint32_t reserved;
  uint64_t probe_blocks;      /* partner blocks of which single records were read (probe path) */
} wsr_batch_stats;
int wsr_batch_get_stats(wsr_batch *b, wsr_batch_stats *s);

/* Cross-shard merge: gathered = n_shards consecutive result arrays as produced
 * by wsr_batch_device_results on every shard (after an all-gather), all DEVICE pointers.
 * Writes the global top-k per query, ordered (score desc, doc id asc). */
int wsr_merge_topk_device(const void *d_gathered_hits, const void *d_gathered_n_hits,
                          int n_shards, int n_queries, int k_stride, void *d_out_hits,
                          void *d_out_n_hits, void *stream);

/* ---- multi-GPU: the shard exchange and document-partitioned groups ----------
 * The reference has one engine per process (SearchEngineServiceNew, engine_services.h:14-27); a
 * document-partitioned deployment of it would run one engine per partition and merge their top-k.
 * Here that lives behind the C ABI: NCCL is called directly from the library (loaded with dlopen
 * on first use: libnccl.so.2, or the path in WSR_NCCL_LIB), on the batch's own stream. */
typedef struct wsr_comm wsr_comm;   /* one rank of the exchange: an NCCL communicator + buffers */
#define WSR_COMM_ID_BYTES 128
/* ncclGetUniqueId: rank 0 creates the id, the host program hands it to the other ranks. */
int wsr_comm_unique_id(char id[WSR_COMM_ID_BYTES]);
wsr_comm *wsr_comm_init_rank(const char id[WSR_COMM_ID_BYTES], int rank, int world, int device);
void wsr_comm_destroy(wsr_comm *c);
/* Enqueued on the batch's stream right behind its search kernels; every rank ends with the merged
 * top-k of all shards (score desc, doc id asc; every query of the batch must use k == k_stride).
 * mode 0, scatter: a grouped send/recv gives rank r every shard's lists of ITS slice of the
 * queries (hits and counts in one NCCL launch), rank r merges that slice, a grouped all-gather
 * distributes the merged slices — (2 - 2/N) n k 16 B per rank and 1/N of the merge work.
 * mode 1, all-gather: every rank receives every shard's lists and merges all queries. */
int wsr_batch_exchange(wsr_batch *b, wsr_comm *c, int mode);
int wsr_batch_exchanged_results(wsr_batch *b, wsr_comm *c, void **d_hits, void **d_n_hits);
int wsr_batch_fetch_exchanged(wsr_batch *b, wsr_comm *c, wsr_hit *hits, int32_t *n_hits); /* D2H + sync */

/* A group = a document-partitioned collection: n_dirs partition directories (standalone vacuum
 * indexes with local doc ids, partition i holding global docs [base_i, base_i + n_i)), spread
 * evenly over n_dev devices of this process (partition i on devices[i / (n_dirs / n_dev)]).
 * Single process (dist == NULL): the library exchanges the collection statistics between the
 * partitions itself (N, average length, per-term df; wsr_index_set_global_stats) and creates one
 * exchange rank per device. Multi-process (one device per process): dist names this process's
 * rank, the world size and the communicator id; the caller then sets the collection statistics on
 * each wsr_group_part() with wsr_index_set_global_stats, because only it can reach the other ranks. */
/* A group serves one batch at a time: calls on the same wsr_group must not overlap (the adapter's
 * request coalescer serialises them). */
typedef struct wsr_group wsr_group;
typedef struct {
  int rank, world;
  char comm_id[WSR_COMM_ID_BYTES];
} wsr_group_dist;
wsr_group *wsr_group_open(const char *const *dirs, int n_dirs, const int *devices, int n_dev,
                          int loader_threads, unsigned flags, const wsr_group_dist *dist, char *err,
                          size_t errlen);
void wsr_group_close(wsr_group *g);
int wsr_group_n_parts(const wsr_group *g);
wsr_index *wsr_group_part(wsr_group *g, int i);   /* borrowed; partition i of this process */
/* The replay driver's inner loop over a group (wsr_search_log_ex for one index): every partition
 * parses, looks up and plans the log text on its GPU, searches, and the per-partition top-k are
 * merged (on the device for partitions sharing one, through the exchange across devices). hits,
 * n_hits, doc_freqs, n_doc_freqs as wsr_search_log_ex. In a multi-process job every rank makes
 * the call; ranks that do not face the client pass hits = n_hits = NULL and only search + exchange. */
int wsr_group_search_log(wsr_group *g, const char *text, size_t len, int k, wsr_hit *hits,
                         int32_t *n_hits, uint32_t *doc_freqs, int32_t *n_doc_freqs, int cap_q,
                         int *n_queries);
/* Device-resident form (benchmarks): plan once, run many times. wsr_group_run enqueues search
 * kernels + merges + exchange on every device (asynchronous); wsr_group_stream is the first
 * device's leading stream (cudaStream_t) for CUDA-event timing. */
int wsr_group_load_log(wsr_group *g, const char *text, size_t len, int k, int *n_queries);
int wsr_group_run(wsr_group *g, int mode);
/* wsr_group_run is pipelined across passes: the exchange of pass i runs on a stream of its own (from
 * a staged copy of the local lists) while the search kernels of pass i+1 run on the leading
 * stream. wsr_group_join makes the leading stream wait for the last exchange (call it before
 * recording an end-of-region event there); wsr_group_fetch and wsr_group_sync join by themselves. */
int wsr_group_join(wsr_group *g);
int wsr_group_sync(wsr_group *g);
int wsr_group_stream(wsr_group *g, void **stream);
int wsr_group_fetch(wsr_group *g, wsr_hit *hits, int32_t *n_hits);
int wsr_group_stats(wsr_group *g, uint64_t *listed_postings, uint64_t *n_postings, uint64_t *hbm_bytes,
                    int64_t *n_docs);

#ifdef __cplusplus
}
#endif
#endif
