// Flat weight file ("FRB2WTS1") written by tools/pack_weights.py; takes the place of the TensorRT engine blob the
// reference deserialises from `engineFile` (/root/reference src/retinaface.cpp:31-55, src/arcface.cpp:45-69).
//   header : char magic[8] = "FRB2WTS1"; int32 version = 1; int32 kind; int32 n_tensors; int32 reserved
//   table  : n_tensors x { char name[96]; int32 dtype (0 = f32, 1 = f16); int32 ndim; int64 dims[4]; int64 offset; int64 nbytes }
//   data   : tensor payloads, each 256-byte aligned, offsets from the start of the file
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace frb {

enum WeightKind { kKindArcfaceIR = 1, kKindArcfaceIRSE = 2, kKindRetinaTrim = 3, kKindRetinaFull = 4 };

struct HostTensor {
    int dtype = 0;
    int ndim = 0;
    int64_t dims[4] = {1, 1, 1, 1};
    const uint8_t* data = nullptr;
    int64_t nbytes = 0;
    int64_t numel() const { return dims[0] * dims[1] * dims[2] * dims[3]; }
};

struct WeightFile {
    int kind = 0;
    std::vector<uint8_t> blob;
    std::map<std::string, HostTensor> tensors;

    const HostTensor& get(const std::string& name, int dtype, int64_t numel) const {
        auto it = tensors.find(name);
        if (it == tensors.end()) throw FileError{FR_EFORMAT, "weight file: tensor '" + name + "' missing"};
        if (it->second.dtype != dtype || it->second.numel() != numel)
            throw FileError{FR_EFORMAT, "weight file: tensor '" + name + "' has the wrong type or size"};
        return it->second;
    }
    bool has(const std::string& name) const { return tensors.count(name) != 0; }
};

inline WeightFile load_weight_file(const char* path) {
    if (!path) throw ArgError{"weights path is null"};
    FILE* f = std::fopen(path, "rb");
    if (!f) throw FileError{FR_ENOENT, std::string("Cant find engine file: ") + path};  // src/retinaface.cpp:53
    WeightFile wf;
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (size < 24) {
        std::fclose(f);
        throw FileError{FR_EFORMAT, "weight file too small"};
    }
    wf.blob.resize(static_cast<size_t>(size));
    const size_t got = std::fread(wf.blob.data(), 1, wf.blob.size(), f);
    std::fclose(f);
    if (got != wf.blob.size()) throw FileError{FR_EFORMAT, "short read on weight file"};
    const uint8_t* p = wf.blob.data();
    if (std::memcmp(p, "FRB2WTS1", 8) != 0) throw FileError{FR_EFORMAT, "not an FRB2WTS1 weight file"};
    int32_t hdr[4];
    std::memcpy(hdr, p + 8, 16);
    if (hdr[0] != 1) throw FileError{FR_EFORMAT, "unsupported weight file version"};
    wf.kind = hdr[1];
    const int n = hdr[2];
    const size_t rec = 96 + 4 + 4 + 32 + 8 + 8;
    if (n < 0 || 24 + rec * static_cast<size_t>(n) > wf.blob.size()) throw FileError{FR_EFORMAT, "corrupt tensor table"};
    for (int i = 0; i < n; ++i) {
        const uint8_t* r = p + 24 + rec * i;
        char name[97];
        std::memcpy(name, r, 96);
        name[96] = 0;
        HostTensor t;
        int32_t dt[2];
        std::memcpy(dt, r + 96, 8);
        t.dtype = dt[0];
        t.ndim = dt[1];
        std::memcpy(t.dims, r + 104, 32);
        int64_t on[2];
        std::memcpy(on, r + 136, 16);
        if (on[0] < 0 || on[1] < 0 || static_cast<size_t>(on[0] + on[1]) > wf.blob.size()) throw FileError{FR_EFORMAT, "tensor out of file bounds"};
        const int64_t esz = t.dtype == 1 ? 2 : 4;
        if (t.numel() * esz != on[1]) throw FileError{FR_EFORMAT, std::string("tensor size mismatch: ") + name};
        t.data = p + on[0];
        t.nbytes = on[1];
        wf.tensors[name] = t;
    }
    return wf;
}

}  // namespace frb
