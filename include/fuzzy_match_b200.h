/*
 * fuzzy_match_b200.h -- C ABI of the B200-native fuzzy-match hot path.
 *
 * Drop-in boundary: the reference has no FFI layer; its hot path is the public method
 * fuzzy::FuzzyMatch::match (reference include/fuzzy/fuzzy_match.hh:59-82, src/fuzzy_match.cc:435-681)
 * over an index built by add_tm/sort (fuzzy_match.hh:52-57, src/suffix_array_index.cc:10-30,
 * src/suffix_array.cc:9-102). This header is what a binding for that path links against:
 *   fm_index_create      <- FuzzyMatch::add_tm(id, Tokens) x N + FuzzyMatch::sort()
 *   fm_match_batch       <- FuzzyMatch::match(Tokens, ...) for a batch of patterns (host buffers)
 *   fm_match_batch_device<- same, device-resident inputs/outputs on a caller stream
 *   fm_match_batch_sharded_device (fm_shard_accept_device + NCCL all-gather + fm_merge_accepted_device)
 *                        <- the same path for a TM sharded by sentence-id range over several GPUs
 * Plain pointers and sizes only; status codes, no exceptions; every function is safe to call from
 * several host threads on one shared index (like the reference's const match()).
 *
 * Word ids: int32, >= 2 for vocabulary words (0 = sentence separator, 1 = unknown, as in reference
 * src/vocab_indexer.cc:10-11). Query ids outside [2, vocab_size) or absent from the TM are "unknown".
 * Sentence ids (s_id) number the KEPT sentences consecutively: empty sentences and sentences longer
 * than max_tokens_in_pattern are dropped exactly as SuffixArrayIndex::add_tm does
 * (src/suffix_array_index.cc:16).
 */
#ifndef FUZZY_MATCH_B200_H
#define FUZZY_MATCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_OK 0
#define FM_ERR_INVALID 1   /* bad argument (see fm_last_error) */
#define FM_ERR_CUDA 2      /* CUDA runtime failure */
#define FM_ERR_NOMEM 3
#define FM_MAX_TOKENS 1023 /* hard cap on max_tokens_in_pattern (reference default: 300) */

typedef struct fm_index fm_index;

/* Arguments of FuzzyMatch::match after the pattern (fuzzy_match.hh:59-82), same meaning/defaults. */
typedef struct fm_params {
  float fuzzy;               /* fuzzy threshold in [0,1] */
  int32_t number_of_matches; /* N; 0 = all */
  int32_t no_perfect;        /* skip cost==0 matches of equal length (Sentence overload only) */
  int32_t min_subseq_length; /* ml */
  float min_subseq_ratio;    /* mr */
  float vocab_idf_penalty;
  float insert_cost, delete_cost, replace_cost; /* EditCosts, include/fuzzy/costs.hh:7-29 */
  float contrastive_factor;
  int32_t contrast_reduce;   /* 0 = ContrastReduce::MEAN, 1 = MAX */
  int32_t contrast_buffer;   /* -1 = number_of_matches */
} fm_params;

/* One FuzzyMatch::Match (fuzzy_match.hh:32-46) minus the id string and the token pointer, which the
 * host adapter fills from its own tables; cost is the edit cost the score was derived from. */
typedef struct fm_match {
  uint32_t s_id;
  float score;
  float penalty;
  int32_t max_subseq;
  int32_t length;
  float cost;
} fm_match;

/* Per-stage device times of the last completed batch on this index (milliseconds, CUDA events on
 * the stream the kernels ran on) and the work it did. Filled only when profiling is enabled. */
typedef struct fm_profile {
  float ms_prepare, ms_search, ms_gather, ms_scan, ms_score, ms_replay, ms_total;
  int64_t n_queries, n_query_tokens, n_slices, n_elements, n_survivors, n_matches;
  int32_t launches; /* kernels launched for the batch */
  int32_t retries;  /* workspace regrowths */
  int64_t n_stage2; /* suffix-array elements that passed the signature test of the gather */
  int64_t n_verified; /* exact coverage counts done for them (repeats of a known pair are skipped) */
  float ms_walk, ms_verify; /* the two kernels of ms_gather: range walk with the signature test / exact coverage */
} fm_profile;

/* Build the device index for one GPU from a CSR translation memory.
 *   sfreq_global / n_sent_global: optional IDF statistics of the whole TM when this index is one
 *   sentence-id shard of it (NULL / 0 = use this TM's own); s_id_base is added to every s_id. */
int fm_index_create(const int32_t* tokens, const int64_t* sent_off, int64_t n_sent, int32_t vocab_size,
                    int32_t max_tokens_in_pattern, const uint32_t* sfreq_global, int64_t n_sent_global,
                    int64_t s_id_base, int device, fm_index** out);
void fm_index_destroy(fm_index* index);
/* Flat on-disk form of a built index (replaces the reference's .fmi Boost archive,
 * src/fuzzy_matcher_binarization.cc:10-51): save writes the HBM-resident arrays as they are, load is
 * read + upload (no sort). The id strings stay with the caller. */
int fm_index_save(const fm_index* index, const char* path);
int fm_index_load(const char* path, int device, fm_index** out);
int64_t fm_index_num_sentences(const fm_index* index); /* kept sentences */
int64_t fm_index_num_suffixes(const fm_index* index);
int32_t fm_index_max_tokens_in_pattern(const fm_index* index); /* FuzzyMatch::max_tokens_in_pattern */
int64_t fm_index_device_bytes(const fm_index* index);
/* kept[s_id] = position of that sentence in the CSR given to fm_index_create (for the id table). */
int fm_index_kept_sources(const fm_index* index, int64_t* kept);
/* word-in-sentence frequencies of this TM (sfreq[vocab_size]), reference src/vocab_indexer.cc:73-90 */
int fm_index_sfreq(const fm_index* index, uint32_t* sfreq);
/* Replace the IDF statistics (e.g. with the all-reduced sfreq of every shard of a sharded TM). */
int fm_index_set_idf_stats(fm_index* index, const uint32_t* sfreq_global, int64_t n_sent_global);
/* Host view of a kept sentence (Match::s / Match::length); valid for the life of the index. */
int fm_index_sentence(const fm_index* index, uint32_t local_s_id, const int32_t** tokens, int32_t* length);

/* Sentence API (reference include/fuzzy/sentence.hh:24-48): the "real" surface form of every token and
 * the penalty tokens (itoks: tags, punctuation, separators...) in the gaps between tokens, as
 * FuzzyMatch::add_tm(id, Sentence, Tokens) (fuzzy_match.hh:53) stores them. real[k] = (real form id << 1) |
 * case_class, two tokens have the same real form iff the values are equal, case_class = first character
 * of the real form is one of "LUMC" (src/edit_distance.cc:55); gaps holds n+1 penalty-token ids (0 = none)
 * per sentence at sent_off[s] + s. sent_off / n_sent are those given to fm_index_create. */
int fm_index_set_real(fm_index* index, const int32_t* real, const int32_t* gaps, const int64_t* sent_off, int64_t n_sent);
/* match(const Sentence& real, const Tokens& pattern, ...) (fuzzy_match.hh:70-82) for a batch: adds the
 * real-token / case / penalty-token terms of _edit_distance (src/edit_distance.cc:19-26,30-31,35-38,53-62).
 * itok_dist[a * n_itok + b] = _edit_distance_char of penalty tokens a and b (include/fuzzy/edit_distance.hxx),
 * with row / column 0 = the lengths; ids are shared with fm_index_set_real. */
int fm_match_batch_real(fm_index* index, const int32_t* q_tokens, const int32_t* q_real, const int32_t* q_gaps,
                        const int64_t* q_off, int64_t n_q, const fm_params* params, const int32_t* itok_dist,
                        int32_t n_itok, int64_t cap, fm_match* out, int32_t* out_count);

/* match() for n_q patterns given as host CSR; out is [n_q * cap], out_count[q] = number of matches
 * the reference would append (only min(count, cap) are stored). */
int fm_match_batch(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q,
                   const fm_params* params, int64_t cap, fm_match* out, int32_t* out_count);

/* match() into result vectors that already hold matches. The reference APPENDS to `matches` (src/fuzzy_match.cc:626-679):
 * entries that are there before the call count against number_of_matches, and the contrastive rerank penalises every
 * candidate against them as well (:634-652), in vector order before the matches it selects itself. prior_sid[prior_off[q]
 * .. prior_off[q+1]) are the sentence ids (of this index) already in the vector of query q; out / out_count receive what
 * the call appends. With empty prior lists this is fm_match_batch. */
int fm_match_batch_prior(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fm_params* params,
                         const uint32_t* prior_sid, const int64_t* prior_off, int64_t cap, fm_match* out, int32_t* out_count);

/* FuzzyMatch::subsequence (fuzzy_match.hh:96-102, src/fuzzy_match.cc:238-365) behind its tokenizer, for a batch of
 * patterns (host CSR): the sub-sequences of a pattern are tried by weight -- length, or summed IDF with idf_weighting --
 * until one occurs in a sentence that no_perfect does not skip; among the first number_of_matches sentences of its
 * suffix-array range the one the reference keeps is returned. The text the reference appends to Match::id (the
 * detokenised sub-sequence) is described by (position, length): the caller owns the token strings. The walk order
 * inside a range follows the word ids, so ids must be assigned in first-seen order like VocabIndexer::addWords
 * (src/vocab_indexer.cc:37-50) for the reference's choice among equally good sentences. */
typedef struct fm_subseq {
  uint32_t s_id;
  float score;
  float cost;
  int32_t position; /* first pattern token of the sub-sequence that located the match */
  int32_t length;   /* its length = Match::max_subseq */
  int32_t found;    /* 1 = match, 0 = none */
} fm_subseq;
int fm_subsequence_batch(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, int32_t number_of_matches,
                         int32_t no_perfect, int32_t min_subseq_length, float min_subseq_ratio, int32_t idf_weighting,
                         fm_subseq* out);

/* Same with device-resident buffers; q_off is int32 here ([n_q+1], device). The batch is enqueued on
 * `stream` (a cudaStream_t) and the call returns once that stream has finished it (the worklist
 * overflow counters are read back; a batch that outgrew the workspace is rerun after regrowing it):
 * on return the results are in d_out/d_out_count. n_query_tokens = q_off[n_q] (known to the caller). */
int fm_match_batch_device(fm_index* index, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                          int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out,
                          int32_t* d_out_count, void* stream);

/* Asynchronous form of the two calls above: submit enqueues the copies and kernels of one batch and
 * returns a ticket; fm_ticket_wait blocks until that batch (not later ones) has finished, reruns it if
 * a worklist overflowed, and frees the ticket (also on error). Several tickets may be in flight on one
 * index -- each owns a workspace -- so the host<->device copies of neighbouring batches overlap the
 * kernels, as the reference's CLI overlaps I/O and matching through its queue of futures
 * (cli/src/FuzzyMatch-cli.cc:112-193). The buffers of a submitted batch must stay valid and untouched until
 * its wait returns; host buffers should be pinned for the copies to be asynchronous. A submitted host
 * batch holds at most 2^18 queries / 2^22 tokens. */
typedef struct fm_ticket fm_ticket;
int fm_match_batch_submit(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q,
                          const fm_params* params, int64_t cap, fm_match* out, int32_t* out_count, fm_ticket** ticket);
int fm_match_batch_device_submit(fm_index* index, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                                 int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out,
                                 int32_t* d_out_count, void* stream, fm_ticket** ticket);
int fm_ticket_wait(fm_ticket* ticket);

/* ---- TM sharded by sentence-id range over several GPUs (BASELINE.json north_star; no counterpart in the
 * reference, which is single-process). Each shard is an fm_index built over its sentence range with the global
 * IDF statistics (fm_index_create: sfreq_global, n_sent_global, s_id_base). Per batch every shard runs the
 * whole pipeline on its own sentences INCLUDING the candidate loop of src/fuzzy_match.cc:567-611 and keeps
 * the records that loop accepts (its bound is never tighter than the global one, so this is a superset of
 * what the global loop accepts in that shard); ONE all-gather moves those records (16 bytes each) in blocks of
 * one size;
 * every rank then replays the union in the reference's candidate order -- bit-identical to one index.
 *
 * One accepted record on the wire (16 bytes). */
typedef struct fm_wire {
  uint32_t s_id;    /* global sentence id */
  uint32_t lm_len;  /* longest n-gram match | sentence length << 16 */
  float cost;
  float rowmin_max;
} fm_wire;
/* Size of one shard's block for n_q queries with room for `capacity` records:
 * int32 header[4] (overflow flags, n_q, capacity, accepted records in total) | int32 off[n_q + 1, padded to 4] |
 * fm_wire rec[capacity] -- the accepted records query after query, each in the order its loop accepted them. */
int64_t fm_wire_block_bytes(int64_t n_q, int64_t capacity);
/* Per-shard half: fills d_block (device, caller-owned, fm_wire_block_bytes) with the records the shard's own
 * candidate loop accepts (the header says how many there were; records beyond the capacity are dropped). Returns
 * after the stream has finished; a workspace overflow of this shard is handled inside (rerun). */
int fm_shard_accept_device(fm_index* shard, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                           int64_t n_query_tokens, const fm_params* params, int64_t capacity, void* d_block, void* stream);
/* Cross-shard half: replays the union of n_shards blocks (ascending s_id order) exactly like the single-index
 * candidate loop (bound heap, top-N). total_capacity = sum of the blocks' capacities. If a shard accepted more
 * records than its block holds, *need_capacity = the largest total of one shard and the results are not written:
 * rerun both halves with blocks of at least that capacity (otherwise 0). `index` only provides the device and a
 * workspace (any shard). Contrastive rerank needs the sentences behind the records: FM_ERR_INVALID here (the NCCL
 * entry points below gather them). */
int fm_merge_accepted_device(fm_index* index, int n_shards, const void* const* d_blocks, int64_t total_capacity,
                             const int32_t* d_q_off, int64_t n_q, const fm_params* params, int64_t cap, fm_match* d_out,
                             int32_t* d_out_count, int64_t* need_capacity, void* stream);

/* The two halves around one NCCL all-gather, one process per GPU: the communicator is NCCL's own
 * (libnccl.so.2 is opened at run time; FM_ERR_INVALID if it is missing). Rank 0 calls fm_comm_unique_id and
 * hands the 128 bytes to every rank by any means (the Python mirror broadcasts them with torch.distributed);
 * fm_comm_create is collective. */
#define FM_COMM_ID_BYTES 128
typedef struct fm_comm fm_comm;
int fm_comm_unique_id(void* id_out);
int fm_comm_create(const void* id, int rank, int world, int device, fm_comm** out);
void fm_comm_destroy(fm_comm* comm);
/* One batch against the sharded TM: every rank passes the same queries (device) and gets the complete result
 * in d_out / d_out_count. Collective; returns after the stream has finished. The block capacity follows the
 * number of records recent batches accepted; a batch that needs more is rerun by all ranks together, as is a
 * batch during which some rank had to regrow its workspace.
 * Contrastive rerank (src/fuzzy_match.cc:613-669) compares the accepted sentences with each other: their tokens
 * (the accepted ones only) are gathered behind the merged replay -- every rank fills the sentences it owns into one
 * slab laid out by the merged lists, one ncclAllReduce sums the slabs -- and every rank reranks on the slab. */
int fm_match_batch_sharded_device(fm_index* shard, fm_comm* comm, const int32_t* d_q_tokens, const int32_t* d_q_off,
                                  int64_t n_q, int64_t n_query_tokens, const fm_params* params, int64_t cap,
                                  fm_match* d_out, int32_t* d_out_count, void* stream);
/* Asynchronous form (see fm_match_batch_device_submit): several batches in flight, each on its own stream, so
 * that the kernels and the all-gathers of neighbouring batches overlap. Every rank must submit and wait in the
 * same order. The buffers stay untouched until fm_ticket_wait returns. */
int fm_match_batch_sharded_submit(fm_index* shard, fm_comm* comm, const int32_t* d_q_tokens, const int32_t* d_q_off,
                                  int64_t n_q, int64_t n_query_tokens, const fm_params* params, int64_t cap,
                                  fm_match* d_out, int32_t* d_out_count, void* stream, fm_ticket** ticket);
/* Bytes received by the last all-gather of this communicator and the record capacity of its blocks. */
int64_t fm_comm_last_gather_bytes(const fm_comm* comm);
int64_t fm_comm_block_capacity(const fm_comm* comm);

int fm_set_profiling(fm_index* index, int enabled);
int fm_get_profile(const fm_index* index, fm_profile* out);
const char* fm_last_error(void);
const char* fm_version(void);

#ifdef __cplusplus
}
#endif
#endif
