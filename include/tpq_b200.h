/*
 * tpq_b200 -- C ABI of the B200-native IVFPQ search path.
 *
 * Drop-in boundary for torchpq.index.IVFPQIndex.search() (reference
 * torchpq/index/IVFPQIndex.py:469-524) and the ops below it.  Every entry point
 * takes plain DEVICE pointers, sizes and a cudaStream_t (passed as void*); the
 * caller (PyTorch host code) owns every buffer, the library never allocates,
 * frees or keeps caller memory, never synchronises the host, and launches only
 * on the stream it is given.  Return value: 0 on success, negative tpq_status
 * on error (tpq_last_error() gives the message for the calling thread).
 *
 * All kernels are hand-written CUDA compiled ahead of time for sm_100a.
 */
#ifndef TPQ_B200_H
#define TPQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TPQ_OK = 0,
  TPQ_ERR_BAD_ARG = -1,      /* mirrors the reference's Python asserts (shape / dtype / range) */
  TPQ_ERR_UNSUPPORTED = -2,  /* valid in the reference but not built here (says which) */
  TPQ_ERR_WORKSPACE = -3,    /* caller workspace too small */
  TPQ_ERR_CUDA = -4          /* a CUDA runtime call failed (message carries cudaGetErrorString) */
} tpq_status;

enum { TPQ_METRIC_EUCLIDEAN = 0, TPQ_METRIC_COSINE = 1 };

/* Index state read by search.  Field-for-field the buffers of the reference's
 * CellContainer / BaseContainer / codecs (container/CellContainer.py:46-81,
 * container/BaseContainer.py:32-38, codec/VQCodec.py:15-17, codec/PQCodec.py:36-46),
 * followed by the scan ("shadow") layout derived from them by tpq_relayout_*. */
typedef struct tpq_index {
  int32_t d_vector;          /* d */
  int32_t n_subvectors;      /* M, M % 4 == 0 (contiguous_size 4, IVFPQIndex.py:41) */
  int32_t n_cells;           /* C */
  int32_t metric;            /* TPQ_METRIC_* (reference `distance`) */
  int64_t capacity;          /* address slots = _address2id.shape[0] */
  const float*   vq_codebook;   /* [d, C]        vq_codec.codebook */
  const float*   pq_codebook;   /* [M, d/M, 256] pq_codec.codebook */
  const uint8_t* storage;       /* [M/4, capacity, 4] _storage (may be NULL once the scan layout exists) */
  const uint8_t* is_empty;      /* [capacity] _is_empty */
  const int64_t* cell_start;    /* [C] _cell_start */
  const int64_t* cell_size;     /* [C] _cell_size */
  const int64_t* address2id;    /* [capacity] _address2id */
  /* ---- scan layout (all NULL / 0 until tpq_relayout_codes has run) ---- */
  int32_t m_pad;                /* M rounded up to a multiple of 32 */
  int32_t shard_rank;           /* this shard owns cells c with c % shard_world == shard_rank */
  int32_t shard_world;
  int32_t reserved0;
  int64_t n_blocks;             /* 32-vector blocks in codes_scan */
  const uint8_t*  codes_scan;   /* [n_blocks, m_pad/16, 32 lanes, 16 B], lane-rotated */
  const uint32_t* block_valid;  /* [n_blocks] bit l set = vector l of the block is live */
  const int32_t*  cell_block_start; /* [C+1] first block of each cell (exclusive prefix; [C] = n_blocks) */
  const float*    pq_codebook_t;    /* [256, m_pad, d/M] transposed PQ codebook (zero rows for padding) */
  const float*    pq_norm_t;        /* [256, m_pad] squared norms of the sub-centroids (0 for cosine) */
  /* ---- residual IVFPQ (pq_use_residual=True, IVFPQIndex.py:48-55,160-170); NULL / 0 otherwise ---- */
  const float*    part2_scan;       /* [C, m_pad/64 groups, 256, 64]: -2 <c_cell,m , p_m,code> - |p_m,code|^2 in LUT layout */
  int32_t         residual;         /* 1 = codes are PQ codes of x - vq_centroid(cell) */
  int32_t         reserved1;
} tpq_index;

/* ------------------------------------------------------------------ misc */
int         tpq_version(void);
const char* tpq_last_error(void);
/* 1 if the library was compiled for the device's architecture (sm_100a) */
int         tpq_device_supported(int device);

/* ------------------------------------------------------------------ query preparation
 * cosine pre-normalisation x / (||x||_2 + 1e-9) per column  (util.py:38-43, IVFPQIndex.py:474-475) */
int tpq_normalize_columns(const float* x_dn, int d, int nq, float* out_dn, void* stream);

/* ------------------------------------------------------------------ coarse probe
 * Replaces metric.negative_squared_l2_distance + fn.Topk + the smart-probing block
 * (metric.py:74-94, fn/Topk.py:43-67, IVFPQIndex.py:485-512).
 *   x_dn [d, nq] (already normalised for cosine), vq_codebook [d, C]
 *   -> probe_sims [nq, n_probe] f32 descending, cells [nq, n_probe] i64,
 *      n_probe_list [nq] i64 (ceil(H * n_probe) when smart != 0, else n_probe).
 * ws: at least tpq_coarse_workspace_bytes(d, nq, C) bytes. */
size_t tpq_coarse_workspace_bytes(int d, int nq, int n_cells);
int tpq_coarse_probe(const float* x_dn, const float* vq_codebook, int d, int nq, int n_cells,
                     int n_probe, int smart, float temperature,
                     float* probe_sims, int64_t* cells, int64_t* n_probe_list,
                     void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ ADC look-up table
 * Replaces PQCodec.precompute_adc (codec/PQCodec.py:62-75; MultiKMeans.py:183-223).
 * -> lut [M, nq, 256] f32, the reference's `precomputed` layout. */
int tpq_build_lut(const float* x_dn, const float* pq_codebook, int d, int M, int nq, int metric,
                  float* lut_mqk, void* stream);

/* ------------------------------------------------------------------ fused list scan + top-k, reference layout
 * Exact drop-in for fn.IVFPQTopk.topk / IVFPQTopkCuda.topk (fn/IVFPQTopk.py:54-104,
 * kernels/IVFPQTopkCuda.py:81-142, kernels/cuda/ivfpq_topk.cu:822-971): same argument
 * list, reads the reference's own storage layout and `precomputed` LUT, accumulates in
 * the reference's order (fp32, m ascending from 0.f).
 *   data [M/4, n_data, 4] u8, precomputed [M, nq, 256] f32, is_empty [n_data] u8,
 *   cell_start/cell_size [nq, n_probe] i64 (already gathered), n_probe_list [nq] i64
 *   -> values [nq, k] f32 descending (-inf padded), address [nq, k] i64 (-1 padded). */
size_t tpq_ivfpq_topk_workspace_bytes(int nq, int k);
int tpq_ivfpq_topk(const uint8_t* data, const float* precomputed, const uint8_t* is_empty,
                   const int64_t* cell_start, const int64_t* cell_size, const int64_t* n_probe_list,
                   int64_t n_data, int M, int nq, int n_probe, int k,
                   float* values, int64_t* address, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ scan layout
 * tpq_relayout_plan: cell_block_start[c] = sum_{c'<c, owned} ceil(cell_extent[c']/32); [C] = total.  The caller passes
 * the cells' CAPACITIES (blocks exist for every slot a cell can hold, so that tpq_store_codes can add items in place;
 * the scan reads only the first ceil(cell_size/32) of them).  Cells not owned by (rank, world) get zero blocks. */
int tpq_relayout_plan(const int64_t* cell_extent, int n_cells, int shard_rank, int shard_world,
                      int32_t* cell_block_start /* [C+1] */, void* stream);
/* Bytes of codes_scan for n_blocks blocks of an M-subvector index. */
size_t tpq_codes_scan_bytes(int M, int64_t n_blocks);
/* Fill codes_scan / block_valid from the reference buffers (index->storage, is_empty,
 * cell_start, cell_size) for the blocks planned in cell_block_start. */
int tpq_relayout_codes(const tpq_index* index, uint8_t* codes_scan, uint32_t* block_valid, void* stream);
/* Residual IVFPQ: the per-cell half of the LUT (IVFPQIndex.precompute_part2, IVFPQIndex.py:160-170) in LUT layout. */
size_t tpq_part2_scan_bytes(int M, int n_cells);
int tpq_relayout_part2(const float* vq_codebook, const float* pq_codebook, int d, int M, int n_cells,
                       float* part2_scan, void* stream);
/* Transposed PQ codebook + norms used by the in-kernel LUT build. */
int tpq_relayout_codebook(const float* pq_codebook, int d, int M, int metric,
                          float* pq_codebook_t, float* pq_norm_t, void* stream);

/* ------------------------------------------------------------------ full search
 * IVFPQIndex.search (IVFPQIndex.py:469-524): x [d, nq] f32 -> values [nq,k] f32 (desc),
 * ids [nq,k] i64, address [nq,k] i64 (nullable).  Needs the scan layout in `index`.
 * keys_out (nullable, [nq,k] u64): packed (score,address) keys for the cross-shard merge. */
size_t tpq_search_workspace_bytes(const tpq_index* index, int nq, int n_probe, int k);
int tpq_ivfpq_search(const tpq_index* index, const float* x_dn, int nq, int n_probe, int k,
                     int smart, float temperature,
                     float* values, int64_t* ids, int64_t* address, uint64_t* keys_out,
                     void* ws, size_t ws_bytes, void* stream);

/* search over given probe lists (IVFPQIndex.search_cells non-residual branch, IVFPQIndex.py:407-467) */
int tpq_ivfpq_search_cells(const tpq_index* index, const float* x_dn, const int64_t* cells,
                           const float* base_sims /* [nq, n_probe], required when index->residual */,
                           const int64_t* n_probe_list, int nq, int n_probe, int k,
                           float* values, int64_t* ids, int64_t* address, uint64_t* keys_out,
                           void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ k-means / codec neighbours of the path
 * tpq_max_sim: MultiKMeans assignment, MaxSimCuda(dim=2, mode "tn") (kernels/MaxSimCuda.py:184-238,
 *   kernels/cuda/max_sim.cu:182-309; caller clustering/MultiKMeans.py:314-333):
 *   data [l, d, n], centroids [l, d, k] -> maxsims [l, n] f32, labels [l, n] i64.
 *   metric EUCLIDEAN: sum_e -(x-c)^2 (direct form, max_sim.cu:78-98); COSINE: sum_e x*c ("inner").
 *   Ties: lowest centroid index (the reference's cross-tile label write is racy, max_sim.cu:173-178).
 * tpq_compute_centroids: ComputeCentroidsCuda (kernels/ComputeCentroidsCuda.py:43-81, kernels/cuda/compute_centroids.cu:9-86):
 *   data [l, d, n], labels [l, n] i64 -> centroids [l, d, k] = member mean, 0 for an empty cluster.
 * tpq_pq_decode: PQDecodeCuda (kernels/PQDecodeCuda.py:38-65, kernels/cuda/pq_decode.cu:7-53):
 *   codebook [M, dsub, 256], code [M, n] u8 -> out [M*dsub, n] f32. */
/* exact == 1: fp32 SIMT kernel, the reference's arithmetic bit for bit.  exact == 0 or 2: TF32 tcgen05 kernel where
 * the shape allows (euclidean, d % 8 == 0, d <= 64, k <= 256, n % 4 == 0), else the exact kernel: labels may differ
 * from the exact ones only between centroids closer than TF32 rounding.  maxsims: exact == 2 -> recomputed exactly
 * in fp32 for the chosen centroid; exact == 0 -> 2*score - |x|^2 from the TF32 score (~1e-3 relative). */
int tpq_max_sim(const float* data, const float* centroids, int l, int d, int64_t n, int k, int metric,
                int exact, float* maxsims, int64_t* labels, void* stream);
size_t tpq_compute_centroids_workspace_bytes(int l, int k);
int tpq_compute_centroids(const float* data, const int64_t* labels, int l, int d, int64_t n, int k,
                          float* centroids, void* ws, size_t ws_bytes, void* stream);
int tpq_pq_decode(const float* codebook, const uint8_t* code, int M, int dsub, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------ build side: placement of new items
 * IVFPQIndex.add -> CellContainer.add (container/CellContainer.py:313-367).
 * tpq_get_ioa: index of appearance of every label among equal labels, input order (kernels/cuda/get_ioa.cu:8-47);
 *   counts [n_cells] (nullable) = items per cell (what `cells.unique(return_counts=True)` feeds `_cell_size[...] +=`).
 * tpq_empty_prefix: prefix[a] = empty slots in [0, a), a = 0..capacity (only needed when the container has holes).
 * tpq_get_write_address: the ioa-th empty slot of the item's cell (kernels/cuda/get_write_address_v2.cu:9-41);
 *   empty_prefix == NULL asserts "no holes": the empties of a cell are [start + size, start + capacity).  -1 = no room.
 * tpq_store_codes: set_data_by_address (CellContainer.py:213-239) + address2id / is_empty / cell_size updates
 *   (:354-360; cell_size becomes max(old, slot + 1)).  codes [M, n] u8.  If codes_scan / block_valid (the scan layout of
 *   `index`, non-const here) are given they are updated in place, so the layout stays valid after the add.
 * tpq_expand_move: CellContainer.expand (:249-311) for many cells at once: every old slot a moves to a + shift[cell(a)];
 *   the caller pre-fills the new buffers (storage 0, address2id -1, is_empty 1) and updates the per-cell tables. */
size_t tpq_ioa_workspace_bytes(int64_t n, int n_cells);
int tpq_get_ioa(const int64_t* cells, int64_t n, int n_cells, int64_t* ioa, int64_t* counts,
                void* ws, size_t ws_bytes, void* stream);
size_t tpq_empty_prefix_workspace_bytes(int64_t capacity);
int tpq_empty_prefix(const uint8_t* is_empty, int64_t capacity, uint32_t* prefix /* [capacity + 1] */,
                     void* ws, size_t ws_bytes, void* stream);
int tpq_get_write_address(const int64_t* cells, const int64_t* ioa, int64_t n,
                          const int64_t* cell_start, const int64_t* cell_size, const int64_t* cell_capacity,
                          int n_cells, const uint32_t* empty_prefix, int64_t* write_adr, void* stream);
int tpq_store_codes(const uint8_t* codes, const int64_t* cells, const int64_t* write_adr, const int64_t* ids,
                    int64_t n, tpq_index* index, uint8_t* storage, int64_t* address2id, uint8_t* is_empty,
                    int64_t* cell_size, uint8_t* codes_scan, uint32_t* block_valid, void* stream);
int tpq_expand_move(const uint8_t* old_storage, const int64_t* old_address2id, const uint8_t* old_is_empty,
                    const int64_t* old_cell_start, const int64_t* shift, int n_cells, int M,
                    int64_t old_capacity, int64_t new_capacity,
                    uint8_t* storage, int64_t* address2id, uint8_t* is_empty, void* stream);

/* ------------------------------------------------------------------ measurement hook (bench.py roofline leg)
 * When enabled, every scan-kernel launch is bracketed by CUDA events on its own stream;
 * tpq_profile_scan_ms waits for them and returns the summed duration and launch count since
 * the last call.  Process-wide, not thread-safe; leave disabled in production. */
int tpq_profile_enable(int on);
int tpq_profile_scan_ms(float* total_ms, int* n_launches);

/* ------------------------------------------------------------------ fused scan + exchange (multi-GPU)
 * The scan of tpq_ivfpq_search(_cells) with its output fused into the cross-shard exchange: every CTA stores its sorted
 * top-k keys straight into slot `my_slot` of each rank's gather buffer peer_keys[p] ([n_slots, nq, k] u64 in peer-mapped
 * memory, e.g. torch symmetric memory) over NVLink, instead of a local buffer that a collective would copy afterwards.
 * The caller then synchronises the ranks (symmetric-memory barrier) and runs tpq_merge_topk on its own buffer.
 * cells == NULL: the coarse probe runs inside (x as given to search); else x / cells / base_sims / n_probe_list as for
 * tpq_ivfpq_search_cells.  TPQ_ERR_UNSUPPORTED when the batch is small enough to be sliced (use the gathered path). */
int tpq_ivfpq_scan_push(const tpq_index* index, const float* x_dn, const int64_t* cells, const float* base_sims,
                        const int64_t* n_probe_list, int nq, int n_probe, int k, int smart, float temperature,
                        uint64_t* const* peer_keys, int n_peers, int my_slot,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ cross-shard merge
 * keys_in [n_parts, nq, k] (the all-gather's layout; each part sorted descending, 0 = empty slot) -> top-k per query,
 * decoded: values (-inf pad), address (-1 pad), ids = address2id[address] (BaseContainer.py:58-65). */
int tpq_merge_topk(const uint64_t* keys_in, int nq, int n_parts, int k,
                   const int64_t* address2id, int64_t capacity,
                   float* values, int64_t* ids, int64_t* address, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPQ_B200_H */
