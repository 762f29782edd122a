// Request batching in front of a pipeline: the serving loop around the body of the reference's /inference handler
// (/root/reference src/app.cpp:293-352). The reference serves one request at a time behind a lock (src/app.cpp:367): one frame per
// detector / embedder call. Here callers on any number of threads hand in ONE frame each and block; a worker thread gathers the frames
// that are waiting (up to the pipeline's max_batch, or whatever arrived within max_wait_us of the oldest one), assembles them in a
// pinned staging batch, runs the batch through fr_pipeline_submit / fr_pipeline_collect with two batches in flight, and hands every
// caller the results of its own frame. JPEG decoding and the HTTP server are not part of this library (out of scope).
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"

using namespace frb;

namespace {

struct Request {
    const uint8_t* frame = nullptr;
    int stride = 0;
    FrBbox* boxes = nullptr;
    int* count = nullptr;
    int64_t* idx = nullptr;
    float* score = nullptr;
    std::chrono::steady_clock::time_point arrived;
    int status = FR_OK;
    std::string error;
    bool done = false;
};

struct Batch {
    std::vector<Request*> reqs;
    int slot = 0;  // staging buffer
};

}  // namespace

struct FrService {
    FrPipeline* pipe = nullptr;
    int device = 0, frame_h = 0, frame_w = 0, max_batch = 0, max_faces = 0;
    std::chrono::microseconds max_wait{200};
    std::mutex mu;
    std::condition_variable q_cv;     // worker: work arrived / stop
    std::condition_variable done_cv;  // callers: some request finished
    std::deque<Request*> queue;
    bool stop = false;
    std::thread worker;
    uint8_t* staging[2] = {nullptr, nullptr};  // pinned, max_batch frames each
    // per-batch result buffers (host), reused
    std::vector<FrBbox> boxes[2];
    std::vector<int> counts[2];
    std::vector<int64_t> idx[2];
    std::vector<float> score[2];
    std::atomic<int64_t> n_batches{0}, n_frames{0};
};

namespace {

void finish_batch(FrService* s, Batch& b, int status, const std::string& err) {
    const int mf = s->max_faces;
    std::lock_guard<std::mutex> lk(s->mu);
    for (size_t i = 0; i < b.reqs.size(); ++i) {
        Request* r = b.reqs[i];
        r->status = status;
        r->error = err;
        if (status == FR_OK) {
            const int c = s->counts[b.slot][i];
            *r->count = c;
            std::copy(s->boxes[b.slot].begin() + i * mf, s->boxes[b.slot].begin() + (i + 1) * mf, r->boxes);
            if (r->idx) std::copy(s->idx[b.slot].begin() + i * mf, s->idx[b.slot].begin() + (i + 1) * mf, r->idx);
            if (r->score) std::copy(s->score[b.slot].begin() + i * mf, s->score[b.slot].begin() + (i + 1) * mf, r->score);
        }
        r->done = true;
    }
    s->done_cv.notify_all();
}

void worker_loop(FrService* s) {
    cudaSetDevice(s->device);
    std::deque<Batch> in_flight;
    int next_slot = 0;
    const size_t row_bytes = static_cast<size_t>(s->frame_w) * 3, frame_bytes = row_bytes * s->frame_h;
    auto collect_one = [&] {
        Batch b = std::move(in_flight.front());
        in_flight.pop_front();
        const int rc = fr_pipeline_collect(s->pipe, s->boxes[b.slot].data(), s->counts[b.slot].data(), s->idx[b.slot].data(),
                                           s->score[b.slot].data(), nullptr);
        finish_batch(s, b, rc, rc == FR_OK ? std::string() : std::string(fr_last_error()));
    };
    for (;;) {
        Batch b;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            if (in_flight.empty()) s->q_cv.wait(lk, [&] { return s->stop || !s->queue.empty(); });
            if (s->stop && s->queue.empty() && in_flight.empty()) return;
            if (!s->queue.empty()) {
                // gather: wait for a full batch only while the GPU still has something to chew on or the oldest request is young
                const auto deadline = s->queue.front()->arrived + s->max_wait;
                if (in_flight.empty())
                    s->q_cv.wait_until(lk, deadline, [&] { return s->stop || static_cast<int>(s->queue.size()) >= s->max_batch; });
                while (!s->queue.empty() && static_cast<int>(b.reqs.size()) < s->max_batch) {
                    b.reqs.push_back(s->queue.front());
                    s->queue.pop_front();
                }
            }
        }
        if (!b.reqs.empty()) {
            // the staging buffer of a slot is free again once the batch that used it was collected: at most two batches are in flight
            if (in_flight.size() == 2) collect_one();
            b.slot = next_slot;
            next_slot ^= 1;
            for (size_t i = 0; i < b.reqs.size(); ++i) {
                const Request* r = b.reqs[i];
                uint8_t* dst = s->staging[b.slot] + i * frame_bytes;
                if (r->stride == static_cast<int>(row_bytes)) {
                    std::memcpy(dst, r->frame, frame_bytes);
                } else {
                    for (int y = 0; y < s->frame_h; ++y) std::memcpy(dst + y * row_bytes, r->frame + static_cast<size_t>(y) * r->stride, row_bytes);
                }
            }
            const int rc = fr_pipeline_submit(s->pipe, s->staging[b.slot], static_cast<int>(row_bytes), static_cast<int>(b.reqs.size()), 0);
            if (rc != FR_OK) {
                finish_batch(s, b, rc, fr_last_error());
            } else {
                s->n_batches.fetch_add(1);
                s->n_frames.fetch_add(static_cast<int64_t>(b.reqs.size()));
                in_flight.push_back(std::move(b));
            }
        } else if (!in_flight.empty()) {
            collect_one();  // nothing new to enqueue: finish the oldest batch
        }
    }
}

}  // namespace

// internal hook (pipeline.cu)
namespace frb {
void pipeline_dims(const FrPipeline* p, int* device, int* frame_h, int* frame_w, int* max_batch, int* max_faces);
}

extern "C" {

int fr_service_create(FrPipeline* p, int max_wait_us, FrService** out) {
    return guarded([&] {
        if (!p || !out) throw ArgError{"null argument"};
        if (max_wait_us < 0) throw ArgError{"max_wait_us must be >= 0"};
        if (fr_pipeline_in_flight(p) != 0) throw StateError{"fr_service_create: the pipeline has batches in flight"};
        std::unique_ptr<FrService> s(new FrService());
        s->pipe = p;
        pipeline_dims(p, &s->device, &s->frame_h, &s->frame_w, &s->max_batch, &s->max_faces);
        s->max_wait = std::chrono::microseconds(max_wait_us);
        DeviceGuard dg(s->device);
        const size_t bytes = static_cast<size_t>(s->max_batch) * s->frame_h * s->frame_w * 3;
        const size_t slots = static_cast<size_t>(s->max_batch) * s->max_faces;
        try {
            for (int k = 0; k < 2; ++k) {
                FRB_CUDA(cudaMallocHost(&s->staging[k], bytes));
                s->boxes[k].resize(slots);
                s->counts[k].resize(s->max_batch);
                s->idx[k].resize(slots);
                s->score[k].resize(slots);
            }
        } catch (...) {
            for (int k = 0; k < 2; ++k) cudaFreeHost(s->staging[k]);
            throw;
        }
        s->worker = std::thread(worker_loop, s.get());
        *out = s.release();
    });
}

void fr_service_destroy(FrService* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->q_cv.notify_all();
    if (s->worker.joinable()) s->worker.join();
    for (int k = 0; k < 2; ++k) cudaFreeHost(s->staging[k]);
    delete s;
}

int fr_service_infer(FrService* s, const uint8_t* frame, int stride, FrBbox* boxes, int* count, int64_t* top1_idx, float* top1_score) {
    return guarded([&] {
        if (!s || !frame || !boxes || !count) throw ArgError{"null argument"};
        if (stride < s->frame_w * 3) throw ArgError{"stride smaller than frame_w * 3"};
        Request r;
        r.frame = frame;
        r.stride = stride;
        r.boxes = boxes;
        r.count = count;
        r.idx = top1_idx;
        r.score = top1_score;
        r.arrived = std::chrono::steady_clock::now();
        {
            std::unique_lock<std::mutex> lk(s->mu);
            if (s->stop) throw StateError{"fr_service_infer: the service is shutting down"};
            s->queue.push_back(&r);
            s->q_cv.notify_one();
            s->done_cv.wait(lk, [&] { return r.done; });
        }
        if (r.status != FR_OK) {
            if (r.status == FR_EINVAL) throw ArgError{r.error};
            if (r.status == FR_ESTATE) throw StateError{r.error};
            throw CudaError{r.error};
        }
    });
}

int fr_service_stats(const FrService* s, int64_t* batches, int64_t* frames) {
    if (!s) return FR_EINVAL;
    if (batches) *batches = s->n_batches.load();
    if (frames) *frames = s->n_frames.load();
    return FR_OK;
}

}  // extern "C"
