/*
 * fr_b200.h — C ABI of the B200-native face-recognition hot path.
 *
 * This is the drop-in boundary underneath the reference's C++ classes. The reference
 * (nghiapq77/face-recognition-cpp-tensorrt) has no FFI layer of its own: its "operator API"
 * is the public part of src/retinaface.h, src/arcface.h, src/matmul.h and src/common.h, compiled
 * into the same executable. The header-only C++ shim in
 * face-recognition-cpp-tensorrt_b200/cpp/{common,retinaface,arcface,matmul}.h re-creates those classes
 * on top of the entry points below; each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FR_E* code on failure; the message for the
 *     calling thread's last failure is returned by fr_last_error(). No C++ exception crosses the ABI.
 *   - handles are opaque, own all their device memory and one CUDA stream, and are NOT re-entrant
 *     (same contract as the reference objects: one instance = one stream, src/retinaface.h:44,
 *     src/arcface.h:55, src/matmul.h:31). Different handles may be used from different threads.
 *   - all output buffers are caller-allocated. "host" pointers are ordinary host memory, "_dev"
 *     entry points take device pointers on the handle's device plus a cudaStream_t passed as void*
 *     (NULL = the handle's own stream) and do not synchronise.
 *   - there is no CPU fallback: if no sm_100 device is present, *_create fails with FR_ENODEVICE.
 */
#ifndef FR_B200_H
#define FR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_OK 0
#define FR_EINVAL (-1)    /* bad argument */
#define FR_ENODEVICE (-2) /* no usable sm_100 GPU */
#define FR_ECUDA (-3)     /* CUDA runtime/driver error (text in fr_last_error) */
#define FR_ENOENT (-4)    /* weight file missing ("Cant find engine file", src/retinaface.cpp:53) */
#define FR_EFORMAT (-5)   /* weight file malformed */
#define FR_ESTATE (-6)    /* call not valid in this state (e.g. empty gallery, src/arcface.cpp:198) */

const char *fr_last_error(void);
/* library/ABI version, bumped on any signature change */
int fr_abi_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t fr_launch_count(void);

/* ---- a1: detection record. Same layout as `struct Bbox`, src/common.h:13-16.
 *      x = row (vertical), y = column (horizontal): src/retinaface.cpp:165,171-174. */
typedef struct FrBbox {
    int x1, y1, x2, y2;
    float score;
} FrBbox;

/* =====================================================================================
 * Gallery / cosine-similarity search  (replaces MatMul, src/matmul.{h,cpp}, and the host
 * argmax ArcFaceIR50::getOutputs, src/arcface.cpp:203-217)
 * ===================================================================================== */
typedef struct FrGallery FrGallery;

/* MatMul::MatMul + MatMul::init (src/matmul.cpp:3-34): upload `rows` (n x dim, row-major f32, host)
 * to `device`. The rows are copied; the caller may free them. dim must be 512 (rec_outputDim,
 * app/config.json:16). `row_offset` is added to every index the search returns (global row id of
 * local row 0 when the gallery is one shard of a row-partitioned gallery; 0 otherwise).
 * n may be 0 (search calls then fail with FR_ESTATE, as featureMatching throws, src/arcface.cpp:198). */
int fr_gallery_create(const float *rows, int64_t n, int dim, int device, int64_t row_offset, FrGallery **out);
/* Same, rows already on `device` (f32). */
int fr_gallery_create_dev(const float *rows_dev, int64_t n, int dim, int device, int64_t row_offset, FrGallery **out);
/* Bench/test utility: fill a gallery on the device with a counter-based generator so that the host
 * can regenerate any row bit-exactly (oracle/search_oracle.py: synth_rows). Row r (global id
 * row_offset + r) = L2-normalised vector of 512 hashed normals keyed by (seed, global id). */
int fr_gallery_create_synthetic(int64_t n, int dim, uint64_t seed, int device, int64_t row_offset, FrGallery **out);
void fr_gallery_destroy(FrGallery *g);
int64_t fr_gallery_rows(const FrGallery *g);
/* CUDA device the gallery (shard) lives on; global row id of its local row 0 */
int fr_gallery_device(const FrGallery *g);
int64_t fr_gallery_row_offset(const FrGallery *g);
/* copy rows [first, first+count) (f32) back to the host — test hook */
int fr_gallery_read_rows(FrGallery *g, int64_t first, int64_t count, float *out_rows);

/* Gallery lifecycle without a full re-upload. The reference re-stages every embedding on the host and re-uploads the
 * whole matrix for each enrolment or /reload (ArcFaceIR50::addEmbedding / initKnownEmbeds / resetEmbeddings / initMatMul,
 * src/arcface.cpp:150-164,233-236; src/app.cpp:131-217,354-365; src/db.cpp:316-346), leaking the previous copies. Here the resident
 * copies (f32 master, fp16 scan copy, e4m3 scan copy when enabled, norm bounds) are maintained incrementally. Not concurrent with
 * searches on the same handle; every call waits for the handle's stream.
 *   reserve  grow the device buffers to hold `capacity` rows (content preserved); append grows geometrically by itself
 *   append   add n rows (host f32, n x 512) after the last row; their local indices are [rows, rows + n)
 *   remove   delete local row `row` by moving the LAST row into its slot (*moved_from = index the moved row had, == row when the
 *            last row itself was deleted); the caller applies the same move to its row -> userId table (classNames)
 *   update   replace rows [first, first + n) in place (re-enrolment of a known face: src/app.cpp:131-217 deletes and re-inserts)
 *   clear    drop all rows, keep the buffers (resetEmbeddings) */
int fr_gallery_reserve(FrGallery *g, int64_t capacity);
int fr_gallery_append(FrGallery *g, const float *rows, int64_t n);
int fr_gallery_update(FrGallery *g, int64_t first, const float *rows, int64_t n);
int fr_gallery_remove(FrGallery *g, int64_t row, int64_t *moved_from);
int fr_gallery_clear(FrGallery *g);
int64_t fr_gallery_capacity(const FrGallery *g);

/* ---- Roster: the reference's row -> userId table (ArcFaceIR50::classNames, src/arcface.h:38-40) kept in step with a row-SHARDED
 * gallery, so that enrolment, deletion and /reload (src/db.cpp:316-346, src/arcface.cpp:150-164,233-236, src/app.cpp:131-217,354-365)
 * are incremental updates of the resident shards. SPMD: every rank owns one shard (a gallery created EMPTY with row_offset =
 * rank << 32) and applies the same sequence of calls; the name table of all shards is replicated, device work touches the local
 * shard only, no communication. Global row id = (shard << 32) | local row — exactly what the search / exchange return.
 *   load         n FACE rows in `SELECT * FROM FACE` (rowid) order: user_ids[i] = USR_ID, blobs[i] = EMBEDDING (512 x f32 LE = 2048
 *                bytes, else FR_EFORMAT); contiguous blocks, shard 0 first; replaces what was resident
 *   add          one row to the least-loaded shard (addEmbedding); *out_id = its global id
 *   remove       one row / every row of a user (move-last-row inside the shard, mirrored in the table)
 *   clear        resetEmbeddings
 *   user         classNames[argmax] (src/arcface.cpp:212); NULL for ids that name no resident row
 * local_shard may be NULL (bookkeeping only). */
typedef struct FrRoster FrRoster;
int fr_roster_create(FrGallery *local_shard, int world, int rank, FrRoster **out);
void fr_roster_destroy(FrRoster *r);
int64_t fr_roster_rows(const FrRoster *r);
int64_t fr_roster_shard_rows(const FrRoster *r, int shard);
int fr_roster_load(FrRoster *r, const char *const *user_ids, const void *const *blobs, const int *blob_bytes, int64_t n);
int fr_roster_add(FrRoster *r, const char *user_id, const float *embedding, int64_t *out_id);
int fr_roster_remove(FrRoster *r, int64_t id);
int fr_roster_remove_user(FrRoster *r, const char *user_id, int64_t *removed);
int fr_roster_clear(FrRoster *r);
const char *fr_roster_user(const FrRoster *r, int64_t id);

/* MatMul::calculate (src/matmul.cpp:36-77): out[i*n_rows + j] = <q_i, row_j>, exact fp32
 * (CUDA_R_32F / CUBLAS_COMPUTE_32F, src/matmul.h:24-25). q: nq x dim host f32; out: nq x n_rows host f32. */
int fr_gallery_sims(FrGallery *g, const float *q, int nq, float *out);
int fr_gallery_sims_dev(FrGallery *g, const float *q_dev, int nq, float *out_dev, void *stream);

/* Fused search: featureMatching + getOutputs (src/arcface.cpp:189-217) without materialising the
 * similarity matrix. For each query the k best rows ordered by (score descending, row index ascending)
 * — for k = 1 this is std::max_element's "first maximum wins" (src/arcface.cpp:210).
 * scores: nq x k f32 (exact fp32 dot products), idx: nq x k int64 (row_offset + local row).
 * If fewer than k rows exist, the tail is filled with score = -inf, idx = -1. 1 <= k <= FR_TOPK_MAX. */
#define FR_TOPK_MAX 8
int fr_gallery_topk(FrGallery *g, const float *q, int nq, int k, float *scores, int64_t *idx);
int fr_gallery_topk_dev(FrGallery *g, const float *q_dev, int nq, int k, float *scores_dev, int64_t *idx_dev, void *stream);

/* Cross-shard merge (the step after the all-gather of per-shard results): parts holds
 * n_parts blocks of nq x k (score, idx) results; writes the nq x k best by (score desc, idx asc). */
int fr_topk_merge_dev(const float *scores_parts_dev, const int64_t *idx_parts_dev, int n_parts, int nq, int k, float *scores_dev,
                      int64_t *idx_dev, int device, void *stream);

/* Test/bench hook: which kernels fr_gallery_topk uses. AUTO = exact SIMT path below 2048 rows (latency-bound sizes),
 * fused tensor-core scan + exact re-score otherwise. Both return exact fp32 scores in the same summation order. */
#define FR_PATH_AUTO 0
#define FR_PATH_EXACT 1
#define FR_PATH_TENSOR 2
int fr_gallery_set_path(FrGallery *g, int path);
/* Precision of the scan copy the fused kernel streams. Scores returned are exact fp32 either way (re-scored from the f32 rows).
 * FR_SCAN_F16 (default): 1 KiB/row. |coarse - exact| <= 1.25e-3 |q| |row| is a deterministic bound, so the result is the exact fp32
 *   top-k whenever rows and queries stay inside fp16's normal range (|component| <= 65504 and row norms >= ~1e-3; always true for
 *   L2-normalised embeddings). Rows outside that range are refused by create / append with FR_ESTATE unless the gallery is switched to
 *   FR_PATH_EXACT.
 * FR_SCAN_F8 (opt-in, L2-normalised rows only): e4m3, 512 B/row, half the HBM traffic and twice the tensor rate. Rows and queries are
 *   rounded STOCHASTICALLY (counter-based dither keyed by (seed, row id, column); FR_F8_SEED sets the seed), which makes the coarse
 *   error of any fixed (query, row) pair a sum of independent bounded zero-mean terms whatever the data look like; every query is
 *   certified after the re-score (best exact score within a gap of the best coarse score) or recomputed by the exact fp32 scan. The
 *   returned top-1 differs from the exact fp32 top-1 with probability <= 1e-12 per query — over the dither, for ARBITRARY rows and
 *   queries chosen without knowledge of the seed (Hoeffding; csrc/search_kernels.cuh kF8LogP; k > 1: <= k * 1e-12).
 *   Query components beyond +-1.75 saturate e4m3: such queries are always recomputed exactly. Top-1 searches (the reference's
 *   getOutputs) use the append epilogue and stay fast for queries without a match; k > 1 on the e4m3 copy keeps the sorted-list
 *   epilogue, whose 16-entry lists overflow under the wide margin when nothing matches — such queries are recomputed by the exact fp32
 *   scan (correct, slow): prefer FR_SCAN_F16 for k > 1. */
#define FR_SCAN_F16 0
#define FR_SCAN_F8 1
int fr_gallery_set_scan(FrGallery *g, int scan);
/* Test hook (parity of the e4m3 image with oracle/f8_dither.py): what = 0: `count` rows of the e4m3 scan copy from row `first`
 * (count x 512 bytes); 1: the query operand image of the last <= 256-query chunk (256 x 512 bytes e4m3 or 256 x 512 fp16);
 * 2: q_margin[256] f32 (accumulator units); 3: q_gap[256] f32; 4: {gmax, g4max, w4max} f32; 5 / 6: the sorted candidate lists of the
 * last k > 1 search, [296 lists][256][16] coarse scores f32 / local rows i32. Waits for the gallery's stream. */
int fr_gallery_debug_read(FrGallery *g, int what, int64_t first, int64_t count, void *out);

/* Cross-GPU exchange of per-shard results over NVLink peer memory, fused into the search (csrc/exchange_impl.cuh; replaces
 * "NCCL all-gather, then fr_topk_merge_dev"). One FrExchange per rank (GPU). Setup: create on every rank, all-gather the
 * fr_exchange_handle_bytes()-byte handles through any host channel, connect. Step, per query batch (<= 256 queries), on one stream:
 *   fr_gallery_topk_push_dev    this shard's search; its re-rank kernel also stores every finished query's results into all peers'
 *                               mailboxes and publishes a flag (no extra launch; an empty shard pushes (-inf, -1))
 *   fr_exchange_wait_merge_dev  waits for all shards' pushes of the OLDEST unmerged batch, merges by (score desc, global row asc)
 * All ranks issue the same sequence of calls. The merge of batch i may be issued after the search of batch i + 1 (at most one batch of
 * lag: four mailbox slots), which hides the wait for the slowest GPU behind the next scan. No host synchronisation, no NCCL: the
 * step is CUDA-graph capturable. A peer that never arrives is not trapped on: after FR_XCHG_TIMEOUT_MS (default 20000) the merge
 * substitutes (-inf, -1) for that shard and fr_exchange_status reports it.
 * fr_exchange_merge_dev is the unfused form for results already in device memory (stand-alone push kernel + wait/merge kernel);
 * local_scores_dev == NULL pushes (-inf, -1). */
typedef struct FrExchange FrExchange;
int fr_exchange_create(int device, int world, int rank, int nq_max, int k_max, FrExchange **out);
int fr_exchange_handle_bytes(void);
int fr_exchange_local_handle(FrExchange *x, void *out_handle);
/* all_handles: world x fr_exchange_handle_bytes() bytes, rank-major (one process per GPU; CUDA IPC) */
int fr_exchange_connect(FrExchange *x, const void *all_handles);
/* same-process groups (single-process multi-GPU hosts, tests): all = the world exchange objects, rank-major */
int fr_exchange_connect_local(FrExchange *x, FrExchange *const *all);
int fr_gallery_topk_push_dev(FrGallery *g, FrExchange *x, const float *q_dev, int nq, int k, float *local_scores_dev,
                             int64_t *local_idx_dev, void *stream);
int fr_exchange_wait_merge_dev(FrExchange *x, int nq, int k, float *scores_dev, int64_t *idx_dev, void *stream);
int fr_exchange_merge_dev(FrExchange *x, const float *local_scores_dev, const int64_t *local_idx_dev, int nq, int k, float *scores_dev,
                          int64_t *idx_dev, void *stream);
/* 0 = every wait so far was satisfied; r + 1 = a merge gave up waiting for rank r. Synchronises the device. */
int fr_exchange_status(FrExchange *x, int *out);
void fr_exchange_destroy(FrExchange *x);
/* The public host-buffer search call of a (possibly sharded) gallery — what each rank's host code calls per batch of <= 256 queries:
 * host queries -> H2D -> this shard's fused search (+ push) -> cross-GPU merge -> D2H of nq x k (score, global row) -> synchronise.
 * x == NULL: single shard (featureMatching + getOutputs, src/arcface.cpp:189-217). stream: cudaStream_t as void*, NULL = the
 * gallery's own stream. With x != NULL every rank must call it with the same queries. */
int fr_search_topk(FrGallery *g, FrExchange *x, const float *q, int nq, int k, float *scores, int64_t *idx, void *stream);

/* Asynchronous form of fr_search_topk for serving: up to THREE batches in flight, so the GPU always has the next search queued while
 * the host collects a previous result (with a lagged cross-GPU merge the result of batch i is only final after the search of batch
 * i + 1: a third batch keeps the search stream fed across the host's collect -> submit turnaround); the H2D of the queries and the
 * D2H of the results run on a second stream beside the scans. On a sharded gallery the cross-GPU merge of batch i is issued after
 * the search of batch i + 1. Queries are copied into library-owned pinned memory inside submit (the caller's buffer is free on
 * return); collect blocks for the OLDEST batch in flight and writes its nq x k results. All ranks of a sharded gallery submit /
 * collect in the same order. submit fails with FR_ESTATE when three batches are already in flight, collect when none is. */
typedef struct FrSearchStream FrSearchStream;
int fr_search_stream_create(FrGallery *g, FrExchange *x, int k, FrSearchStream **out);
void fr_search_stream_destroy(FrSearchStream *s);
void *fr_search_stream_cuda_stream(FrSearchStream *s);
int fr_search_stream_submit(FrSearchStream *s, const float *q, int nq);
int fr_search_stream_collect(FrSearchStream *s, float *scores, int64_t *idx, int *nq_out);

/* roofline bookkeeping for bench.py: algorithmic bytes/flops of the dominant kernel of the last topk call */
typedef struct FrSearchStats {
    int64_t scan_bytes; /* bytes of the resident scan copy streamed (n_rows * dim * 2 for the fp16 copy, * 1 for the e4m3 copy) */
    int64_t flops;      /* 2 * nq_padded * n_rows * dim */
    int launches;       /* kernels launched by the last call */
    int ctas;           /* CTAs of the fused kernel */
} FrSearchStats;
int fr_gallery_last_stats(const FrGallery *g, FrSearchStats *out);
/* diagnostic: how many queries of the last fr_gallery_topk* call on the tensor path the re-rank handed to the exact fp32 scan
 * (candidate set possibly incomplete). Waits for the gallery's own stream; after a *_dev call on a caller stream, synchronise
 * that stream first. 0 in the normal case. */
int fr_gallery_last_flagged(FrGallery *g, int *out);
/* bench.py's live roofline measurement: when enabled, every launch of the fused scan kernel is bracketed by CUDA events on
 * the stream it is launched on; fr_gallery_scan_time waits for them, returns their summed duration and count, and resets. */
int fr_gallery_set_timing(FrGallery *g, int enable);
int fr_gallery_scan_time(FrGallery *g, double *total_ms, int *launches);
/* enable == 2: one FIXED event pair per scan copy instead of a pool. Fixed events survive stream capture (a replayed CUDA graph
 * re-records them), so the duration of the fused kernel's most recent launch — eager or inside a graph replay — can be read from
 * the very steps that are being timed. Waits for that launch. */
int fr_gallery_last_scan_ms(FrGallery *g, int scan, double *ms);
/* Summed duration of the first `count` event pairs of the pool WITHOUT resetting it. A block of `count` steps captured into one CUDA
 * graph while timing mode 1 is on owns pairs 0 .. count-1: after every replay of that graph this returns the time its `count`
 * launches of the fused kernel took — the same steps the caller times around the replay. */
int fr_gallery_pool_time(FrGallery *g, int count, double *total_ms);

/* =====================================================================================
 * Embedder  (replaces ArcFaceIR50's network half, src/arcface.{h,cpp})
 * ===================================================================================== */
typedef struct FrEmbedder FrEmbedder;
#define FR_MODE_IR 0    /* IR_50    - the deployed model, conversion/arcface/torch2trt.py:4,21 */
#define FR_MODE_IR_SE 1 /* IR_SE_50 - conversion/arcface/model_irse.py:217-222 */

/* ArcFaceIR50::ArcFaceIR50 (src/arcface.cpp:21-43). weights_path takes the place of `engineFile`: a flat weight file written
 * by tools/pack_weights.py; the mode (IR / IR_SE) is read from its header. A missing file fails with FR_ENOENT and the
 * message "Cant find engine file" (src/arcface.cpp:67). max_batch = rec_maxBatchSize (any batch up to it may be run). */
int fr_embedder_create(const char *weights_path, int max_batch, int device, FrEmbedder **out);
void fr_embedder_destroy(FrEmbedder *e);
int fr_embedder_mode(const FrEmbedder *e);
int fr_embedder_max_batch(const FrEmbedder *e);

/* ArcFaceIR50::doInference(float*, float*, int) (src/arcface.cpp:138-148): input batch x 3 x 112 x 112 f32 planar RGB
 * normalised ((x-127.5)*0.0078125), output batch x 512 f32, unit L2 norm. Host buffers. */
int fr_embedder_run(FrEmbedder *e, const float *chw, int batch, float *out512);
/* preprocessFaces + doInference (src/arcface.cpp:116-129,138-148): aligned 112x112 u8 BGR HWC crops in (the /recognize path,
 * src/app.cpp:243-287), embeddings out. Host buffers. */
int fr_embedder_run_crops(FrEmbedder *e, const uint8_t *crops_bgr_u8, int batch, float *out512);
/* device buffers, ordered after/before `stream` (cudaStream_t as void*, NULL = caller synchronises itself) */
int fr_embedder_run_dev(FrEmbedder *e, const float *chw_dev, int batch, float *out512_dev, void *stream);
/* per-layer trace hook for parity debugging: re-runs the last input up to `layer` (0 = input_layer, 1..24 = body units) and
 * copies that activation as f32 NCHW into out (host, cap floats); *n_written = element count. */
int fr_embedder_trace(FrEmbedder *e, int layer, float *out, int64_t cap, int64_t *n_written);

/* =====================================================================================
 * Detector  (replaces RetinaFace, src/retinaface.{h,cpp})
 * ===================================================================================== */
typedef struct FrDetector FrDetector;

/* RetinaFace::RetinaFace (src/retinaface.cpp:3-29). weights_path takes the place of `engineFile`: a flat weight file written
 * by tools/pack_weights.py (FR_ENOENT + "Cant find engine file" when missing, src/retinaface.cpp:53). net_h / net_w =
 * det_inputShape[1..2] (multiples of 32); frame_h / frame_w = input_frameHeight / Width; max_faces = det_maxFacesPerScene;
 * nms_thr / bbox_thr = det_threshold_nms / det_threshold_bbox. with_landmarks != 0 also evaluates the landmark head of
 * conversion/retina/models/retinaface.py:37-46 (absent from the deployed trimmed model; needs a "full" checkpoint). */
int fr_detector_create(const char *weights_path, int net_h, int net_w, int frame_h, int frame_w, int max_batch, int max_faces,
                       float nms_thr, float bbox_thr, int with_landmarks, int device, FrDetector **out);
void fr_detector_destroy(FrDetector *d);
/* m_OUTPUT_SIZE_BASE (src/retinaface.cpp:13): anchors per image */
int fr_detector_num_anchors(const FrDetector *d);

/* RetinaFace::findFace (src/retinaface.cpp:147-152) for `batch` frames: preprocess (:106-136), network (:138-145),
 * decode + threshold + sort + NMS + cap (:154-208, :248-271). frames: batch images, u8 HWC BGR, frame_h x frame_w, `stride`
 * bytes per row, images contiguous (image b starts at b * frame_h * stride); host memory. boxes: batch x max_faces (Bbox
 * semantics: x = row, y = column); counts: batch. landmarks (optional, may be NULL): batch x max_faces x 10 floats, (column,
 * row) pairs in frame pixels — an extension, the reference decodes no landmarks. */
int fr_detector_run(FrDetector *d, const uint8_t *frames, int stride, int batch, FrBbox *boxes, int *counts, float *landmarks);
/* Parity hook at the network boundary (the tensors RetinaFace::doInference returns, src/retinaface.cpp:138-145):
 * loc: batch x A x 4, conf: batch x A x 2 (softmax), landm: batch x A x 10 or NULL. Host memory. */
int fr_detector_raw(FrDetector *d, const uint8_t *frames, int stride, int batch, float *loc, float *conf, float *landm);
/* Parity hook for the network alone: input already preprocessed, f32 planar CHW (B,G,R planes, mean subtracted), exactly the
 * tensor RetinaFace::preprocess produces (src/retinaface.cpp:128-135). */
int fr_detector_net(FrDetector *d, const float *chw, int batch, float *loc, float *conf, float *landm);
/* Parity hook for decode + NMS alone (RetinaFace::postprocessing, src/retinaface.cpp:154-208) on caller-provided head
 * outputs (host). */
int fr_detector_post(FrDetector *d, const float *loc, const float *conf, const float *landm, int batch, FrBbox *boxes, int *counts,
                     float *landmarks);
/* device-resident variant used by the end-to-end pipeline: frames and outputs on the detector's device, launched on `stream`
 * (cudaStream_t as void*, NULL = the handle's own stream); does not synchronise. */
int fr_detector_run_dev(FrDetector *d, const uint8_t *frames_dev, int stride, int batch, FrBbox *boxes_dev, int *counts_dev,
                        float *landmarks_dev, void *stream);

/* getCroppedFaces + preprocessFaces + doInference (src/arcface.cpp:3-17,116-129,131-148) for caller-provided boxes of ONE frame
 * (= ArcFaceIR50::forward, src/arcface.cpp:166-187): frame u8 HWC BGR (frame_h x frame_w, stride bytes/row), n boxes with Bbox
 * semantics (ROI = Rect(Point(y1,x1), Point(y2,x2)) = rows [x1,x2), columns [y1,y2)), bicubic resize to 112x112 on the GPU,
 * BGR->RGB, normalise, embed. Row i of out512 = embedding of box i (the reference's chunk-offset bug, src/arcface.cpp:184, is not
 * reproduced). crops_u8 (optional): n x 112 x 112 x 3 BGR u8 crops (= CroppedFace::face). Host buffers. */
int fr_embedder_run_boxes(FrEmbedder *e, const uint8_t *frame, int frame_h, int frame_w, int stride, const FrBbox *boxes, int n,
                          float *out512, uint8_t *crops_u8);

/* =====================================================================================
 * End-to-end pipeline: detect -> crop/resize -> embed -> search on one GPU
 * (the body of the /inference handler, src/app.cpp:293-352, without the socket/JPEG glue)
 * ===================================================================================== */
typedef struct FrPipeline FrPipeline;
/* The pipeline borrows the three handles (they must outlive it and live on the same device); gal may be NULL (no matching). */
int fr_pipeline_create(FrDetector *det, FrEmbedder *emb, FrGallery *gal, FrPipeline **out);
void fr_pipeline_destroy(FrPipeline *p);
/* frames: batch x frame_h x frame_w x 3 u8 BGR host (pinned or pageable), `stride` bytes per row. Outputs per frame f and face
 * slot j < max_faces: boxes[f*max_faces+j] (valid for j < counts[f]), top1_idx / top1_score [f*max_faces+j] (idx = -1 for empty
 * slots or when there is no gallery), embeddings (optional) [f*max_faces+j][512]. */
int fr_pipeline_run(FrPipeline *p, const uint8_t *frames, int stride, int batch, FrBbox *boxes, int *counts, int64_t *top1_idx,
                    float *top1_score, float *embeddings);
/* The same with two batches in flight: submit enqueues a batch (H2D in sub-batches on a copy stream, detection, device-side face-list
 * compaction, crop + embed + search for as many faces as the previous batch had) and returns without waiting for the GPU; collect waits
 * for the OLDEST submitted batch and fills the outputs of fr_pipeline_run. While batch i computes, batch i + 1's frames cross PCIe and
 * the host prepares batch i + 2: the request-batching loop a server runs (src/app.cpp:293-352 handles one request at a time under a
 * lock, :367). `frames` must stay valid and unchanged until the batch is collected. At most two batches may be in flight (FR_ESTATE
 * otherwise); fr_pipeline_run requires none. embeddings may only be collected from a batch submitted with want_embeddings != 0. */
int fr_pipeline_submit(FrPipeline *p, const uint8_t *frames, int stride, int batch, int want_embeddings);
int fr_pipeline_collect(FrPipeline *p, FrBbox *boxes, int *counts, int64_t *top1_idx, float *top1_score, float *embeddings);
int fr_pipeline_in_flight(const FrPipeline *p);

/* Request batching in front of a pipeline (the serving loop around src/app.cpp:293-352; the reference serves one request
 * at a time behind a lock, src/app.cpp:367). fr_service_infer is thread-safe and blocking: callers on any number of threads hand in ONE
 * frame each (frame_h x frame_w x 3 u8 BGR, `stride` bytes per row; it must stay valid until the call returns); a worker thread gathers
 * the frames that are waiting - up to the detector's max_batch, or whatever arrived within max_wait_us of the oldest one while the GPU
 * is idle - into a pinned staging batch, runs it through fr_pipeline_submit / fr_pipeline_collect (two batches in flight) and hands
 * every caller the results of its own frame: boxes[max_faces] (valid for j < *count), top1_idx / top1_score [max_faces] (optional).
 * The service owns the pipeline's submit / collect ring while it lives: do not call fr_pipeline_* on the same pipeline concurrently.
 * JPEG decoding and the HTTP layer stay with the host application. */
typedef struct FrService FrService;
int fr_service_create(FrPipeline *p, int max_wait_us, FrService **out);
void fr_service_destroy(FrService *s);
int fr_service_infer(FrService *s, const uint8_t *frame, int stride, FrBbox *boxes, int *count, int64_t *top1_idx, float *top1_score);
/* monitoring: batches formed and frames served so far */
int fr_service_stats(const FrService *s, int64_t *batches, int64_t *frames);

/* JPEG decode for the serving loop: replaces cv::imdecode + cv::resize of the request handlers (src/app.cpp:247-256
 * "/recognize", :294-301 "/inference"). Huffman decoding on the host, IDCT and colour conversion on the GPU (nvJPEG, loaded on first
 * use), output u8 BGR interleaved (what cv::imdecode returns for a colour JPEG) stretched to out_w x out_h with OpenCV's INTER_LINEAR
 * fixed-point arithmetic when the decoded size differs (out_w <= 0 or out_h <= 0: as decoded, see fr_jpeg_info). `bgr` is a host
 * buffer of out_h rows of `stride` bytes - ready for fr_service_infer / fr_pipeline_run / fr_detector_run. A stream that is not a
 * decodable JPEG fails with FR_EINVAL and the reference's message "Empty image" (src/app.cpp:251,298); a missing nvJPEG library with
 * FR_ENOENT. One handle = one CUDA stream, not re-entrant (one decoder per serving thread). */
typedef struct FrJpegDecoder FrJpegDecoder;
int fr_jpeg_decoder_create(int device, FrJpegDecoder **out);
void fr_jpeg_decoder_destroy(FrJpegDecoder *d);
int fr_jpeg_info(FrJpegDecoder *d, const uint8_t *jpeg, size_t nbytes, int *width, int *height);
int fr_jpeg_decode(FrJpegDecoder *d, const uint8_t *jpeg, size_t nbytes, int out_w, int out_h, uint8_t *bgr, int stride);

#ifdef __cplusplus
}
#endif
#endif /* FR_B200_H */
