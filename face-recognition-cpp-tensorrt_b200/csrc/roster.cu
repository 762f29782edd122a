// Gallery lifecycle around the search path: the row -> userId table of the reference (ArcFaceIR50::classNames,
// /root/reference src/arcface.h:38-40, filled by Database::getEmbeddings src/db.cpp:316-346 and addEmbedding src/arcface.cpp:150-160,
// dropped by resetEmbeddings :233-236) kept in step with a row-SHARDED gallery, so that enrolment, deletion and /reload
// (src/app.cpp:131-217,354-365) become incremental updates of the resident shards instead of "re-read everything, re-upload everything".
//
// SPMD by construction: every rank (one process per GPU, or several roster objects in one process) owns ONE shard and applies the SAME
// sequence of operations; the name table of ALL shards is replicated (it is small: one string per row), the device work only touches
// the local shard. No communication is needed because every decision (which shard takes a new row, which row moves on a delete)
// is a pure function of the replicated state.
//   global row id = (shard << 32) | local row        (the shard's gallery is created with row_offset = shard << 32, so the ids that
//                                                     fr_gallery_topk* / the exchange return resolve here without translation)
//   load     FACE rows in `SELECT * FROM FACE` order (rowid order) are split into contiguous blocks, shard 0 first: with the
//            (score desc, global id asc) merge an exact tie is won by the earlier database row, like the reference's first maximum
//   add      goes to the least-loaded shard (lowest shard on a tie), at its end
//   remove   fr_gallery_remove's move-last-row, mirrored in the name table. After incremental deletes the physical order differs from
//            the database order; only exact score ties between different rows can observe that, and clear + load restores it
// EMBEDDING BLOBs are rec_outputDim = 512 little-endian f32 (the reference stores the raw float array, src/db.cpp:236-262,336).
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

using namespace frb;

struct FrRoster {
    FrGallery* local = nullptr;  // may be null: bookkeeping only (tests of the host logic, ranks without a device)
    int world = 1, rank = 0;
    std::vector<std::vector<std::string>> names;  // [shard][local row] -> userId
};

namespace {
constexpr int kDimR = 512;
constexpr int64_t kShardShift = 32;

void check_roster(const FrRoster* r) {
    if (!r) throw ArgError{"null roster"};
}
int64_t global_id(int shard, int64_t local) { return (static_cast<int64_t>(shard) << kShardShift) | local; }
void split_id(const FrRoster* r, int64_t id, int* shard, int64_t* local) {
    if (id < 0) throw ArgError{"negative row id"};
    *shard = static_cast<int>(id >> kShardShift);
    *local = id & ((int64_t(1) << kShardShift) - 1);
    if (*shard >= r->world || *local >= static_cast<int64_t>(r->names[*shard].size())) throw ArgError{"row id does not name a resident row"};
}
void gcheck(int rc) {
    if (rc != FR_OK) throw StateError{std::string("gallery: ") + fr_last_error()};
}
// delete (shard, local): the shard's last row moves into the slot
void remove_at(FrRoster* r, int shard, int64_t local) {
    auto& v = r->names[shard];
    const int64_t last = static_cast<int64_t>(v.size()) - 1;
    if (shard == r->rank && r->local) {
        int64_t moved = -1;
        gcheck(fr_gallery_remove(r->local, local, &moved));
        if (moved != last) throw StateError{"gallery and roster disagree about the shard's last row"};
    }
    if (local != last) v[local] = std::move(v[last]);
    v.pop_back();
}
}  // namespace

extern "C" {

int fr_roster_create(FrGallery* local_shard, int world, int rank, FrRoster** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (world < 1 || world > 1024 || rank < 0 || rank >= world) throw ArgError{"bad world / rank"};
        if (local_shard) {
            if (fr_gallery_rows(local_shard) != 0) throw StateError{"the roster takes over an EMPTY shard (rows are added through it)"};
            if (fr_gallery_row_offset(local_shard) != global_id(rank, 0))
                throw ArgError{"the shard's gallery must be created with row_offset = rank << 32"};
        }
        std::unique_ptr<FrRoster> r(new FrRoster());
        r->local = local_shard;
        r->world = world;
        r->rank = rank;
        r->names.resize(world);
        *out = r.release();
    });
}

void fr_roster_destroy(FrRoster* r) { delete r; }

int64_t fr_roster_rows(const FrRoster* r) {
    if (!r) return -1;
    int64_t n = 0;
    for (const auto& v : r->names) n += static_cast<int64_t>(v.size());
    return n;
}

int64_t fr_roster_shard_rows(const FrRoster* r, int shard) {
    if (!r || shard < 0 || shard >= r->world) return -1;
    return static_cast<int64_t>(r->names[shard].size());
}

/* Database::getEmbeddings (src/db.cpp:316-346): n FACE rows in `SELECT * FROM FACE` order. user_ids[i] = column USR_ID, blobs[i] /
 * blob_bytes[i] = column EMBEDDING. Replaces whatever was resident (the reference resets before it reloads, src/app.cpp:356-361). */
int fr_roster_load(FrRoster* r, const char* const* user_ids, const void* const* blobs, const int* blob_bytes, int64_t n) {
    return guarded([&] {
        check_roster(r);
        if (n < 0 || (n > 0 && (!user_ids || !blobs || !blob_bytes))) throw ArgError{"bad arguments"};
        for (int64_t i = 0; i < n; ++i) {
            if (!user_ids[i] || !blobs[i]) throw ArgError{"null user id / blob"};
            if (blob_bytes[i] != kDimR * static_cast<int>(sizeof(float)))
                throw FileError{FR_EFORMAT, "EMBEDDING blob of row " + std::to_string(i) + " has " + std::to_string(blob_bytes[i]) +
                                                " bytes, expected 2048 (512 x f32)"};
        }
        if (r->local) gcheck(fr_gallery_clear(r->local));
        for (auto& v : r->names) v.clear();
        const int64_t per = (n + r->world - 1) / r->world;  // contiguous row blocks, shard 0 first
        for (int g = 0; g < r->world; ++g) {
            const int64_t lo = std::min(n, g * per), hi = std::min(n, lo + per);
            for (int64_t i = lo; i < hi; ++i) r->names[g].emplace_back(user_ids[i]);
            if (g == r->rank && r->local && hi > lo) {
                std::vector<float> rows(static_cast<size_t>(hi - lo) * kDimR);
                for (int64_t i = lo; i < hi; ++i) std::memcpy(rows.data() + (i - lo) * kDimR, blobs[i], sizeof(float) * kDimR);  // LE f32, as stored
                gcheck(fr_gallery_append(r->local, rows.data(), hi - lo));
            }
        }
    });
}

/* ArcFaceIR50::addEmbedding (src/arcface.cpp:150-160) without the re-upload: one new row. *out_id = its global row id. */
int fr_roster_add(FrRoster* r, const char* user_id, const float* embedding, int64_t* out_id) {
    return guarded([&] {
        check_roster(r);
        if (!user_id || !embedding) throw ArgError{"null user id / embedding"};
        int target = 0;
        for (int g = 1; g < r->world; ++g)
            if (r->names[g].size() < r->names[target].size()) target = g;
        if (target == r->rank && r->local) gcheck(fr_gallery_append(r->local, embedding, 1));
        r->names[target].emplace_back(user_id);
        if (out_id) *out_id = global_id(target, static_cast<int64_t>(r->names[target].size()) - 1);
    });
}

int fr_roster_remove(FrRoster* r, int64_t id) {
    return guarded([&] {
        check_roster(r);
        int shard;
        int64_t local;
        split_id(r, id, &shard, &local);
        remove_at(r, shard, local);
    });
}

/* every face of one user (the /delete flow removes the user's FACE rows, then reloads). *removed = number of rows dropped. */
int fr_roster_remove_user(FrRoster* r, const char* user_id, int64_t* removed) {
    return guarded([&] {
        check_roster(r);
        if (!user_id) throw ArgError{"null user id"};
        int64_t n = 0;
        for (int g = 0; g < r->world; ++g)
            for (int64_t l = static_cast<int64_t>(r->names[g].size()) - 1; l >= 0; --l)  // descending: a move never skips an unvisited row
                if (r->names[g][l] == user_id) {
                    remove_at(r, g, l);
                    ++n;
                }
        if (removed) *removed = n;
    });
}

/* ArcFaceIR50::resetEmbeddings (src/arcface.cpp:233-236) */
int fr_roster_clear(FrRoster* r) {
    return guarded([&] {
        check_roster(r);
        if (r->local) gcheck(fr_gallery_clear(r->local));
        for (auto& v : r->names) v.clear();
    });
}

/* classNames[argmax] (src/arcface.cpp:212): the userId of a global row id returned by the search. The pointer stays valid until the
 * next mutating call. NULL (and fr_last_error) when the id names no resident row (e.g. -1 = "no rows"). */
const char* fr_roster_user(const FrRoster* r, int64_t id) {
    const char* res = nullptr;
    guarded([&] {
        check_roster(r);
        int shard;
        int64_t local;
        split_id(r, id, &shard, &local);
        res = r->names[shard][local].c_str();
    });
    return res;
}

}  // extern "C"
