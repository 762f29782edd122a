// Drop-in replacement for the reference's src/common.h (/root/reference/src/common.h:13-53, src/common.cpp:3-48): same names,
// no TensorRT / cuBLAS headers needed. Header-only; link libfr_b200.so.
#ifndef COMMON_H
#define COMMON_H

#include <dirent.h>

#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fr_b200.h"

// detection record, same layout as FrBbox: x = row (vertical), y = column (horizontal)   (src/common.h:13-16)
struct Bbox {
    int x1, y1, x2, y2;
    float score;
};
static_assert(sizeof(Bbox) == sizeof(FrBbox), "Bbox must be layout-compatible with FrBbox");

struct Paths {
    std::string absPath;
    std::string className;
};

inline bool fileExists(const std::string &name) {
    std::ifstream f(name.c_str());
    return f.good();
}

// root/<class>/<file>.jpg walker used by the gallery-generation mode of the app (src/common.cpp:8-41)
inline void getFilePaths(std::string rootPath, std::vector<struct Paths> &paths) {
    const std::string postfix = ".jpg";
    DIR *dir = opendir(rootPath.c_str());
    if (!dir) return;
    while (struct dirent *entry = readdir(dir)) {
        const std::string classPath = rootPath + "/" + entry->d_name;
        DIR *classDir = opendir(classPath.c_str());
        if (!classDir) continue;
        while (struct dirent *fileEntry = readdir(classDir)) {
            const std::string name(fileEntry->d_name);
            if (fileEntry->d_type == DT_DIR || name.size() < postfix.size()) continue;
            if (name.compare(name.size() - postfix.size(), postfix.size(), postfix) != 0) continue;
            Paths p;
            p.className = entry->d_name;
            p.absPath = classPath + "/" + name;
            paths.push_back(p);
        }
        closedir(classDir);
    }
    closedir(dir);
}

// The reference's TRTLogger derives from nvinfer1::ILogger (src/common.h:28-53). The applications only default-construct it and
// pass it to the two constructors (src/app.cpp:28,52,56); there is no TensorRT here, so it is an empty tag type.
class TRTLogger {
  public:
    enum class Severity { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    void log(Severity severity, const char *msg) noexcept {
        static const char *names[] = {"INTERNAL_ERROR: ", "ERROR: ", "WARNING: ", "INFO: ", "VERBOSE: "};
        const int s = static_cast<int>(severity);
        std::cerr << (s >= 0 && s <= 4 ? names[s] : "UNKNOWN: ") << msg << std::endl;
    }
};

// C-ABI status -> the exceptions the reference throws (src/common.cpp:43-48, src/retinaface.cpp:53, src/arcface.cpp:67,198)
inline void frCheck(int rc) {
    if (rc == FR_OK) return;
    const std::string msg = fr_last_error();
    if (rc == FR_ENOENT) throw std::logic_error("Cant find engine file");
    if (rc == FR_ESTATE) throw "Feature matching: No faces in database or no faces found";  // caught as const char* (src/app.cpp:276,341)
    std::cerr << "CUDA API failed: " << msg << std::endl;
    throw std::logic_error("CUDA API failed");
}

#endif  // COMMON_H
