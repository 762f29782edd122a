// Drop-in test: the call pattern of the reference's application (/root/reference/src/app.cpp:28-61,89-93,168-170,260-270,
// 304-310,357-361 and src/db.cpp:326-339) compiled against the shim headers of face-recognition-cpp-tensorrt_b200/cpp/.
// Usage: dropin_usage <det.frw> <arc.frw>   -> prints "DROPIN OK" and a few numbers; exit 0.
// With no arguments it only checks the error conventions that need no GPU.
#include <cmath>
#include <cstdio>

#include "arcface.h"
#include "retinaface.h"

int main(int argc, char **argv) {
    TRTLogger gLogger;  // src/app.cpp:28
    const int W = 640, H = 640, maxFaces = 4;
    std::vector<int> detShape = {3, 640, 640}, recShape = {3, 112, 112};
    // missing engine file -> std::logic_error("Cant find engine file") (src/retinaface.cpp:53, src/arcface.cpp:67)
    try {
        RetinaFace bad(gLogger, "/nonexistent/det.engine", W, H, "input_det", {"output_det0", "output_det1"}, detShape, 1, maxFaces, 0.4f, 0.6f);
        return 10;
    } catch (const std::logic_error &e) {
        if (std::string(e.what()) != "Cant find engine file") return 11;
    }
    try {
        ArcFaceIR50 bad(gLogger, "/nonexistent/rec.engine", W, H, "input", "output", recShape, 512, 1, maxFaces, 0.65f);
        return 12;
    } catch (const std::logic_error &e) {
        if (std::string(e.what()) != "Cant find engine file") return 13;
    }
    if (argc < 3) {
        std::printf("DROPIN OK (no-GPU checks only)\n");
        return 0;
    }
    ArcFaceIR50 recognizer(gLogger, argv[2], W, H, "input", "output", recShape, 512, 1, maxFaces, 0.65f);  // src/app.cpp:52
    RetinaFace detector(gLogger, argv[1], W, H, "input_det", {"output_det0", "output_det1"}, detShape, 1, maxFaces, 0.4f, 0.6f);
    std::vector<struct Bbox> outputBbox;
    outputBbox.reserve(maxFaces);

    // featureMatching with an empty gallery throws a const char* the handlers catch (src/arcface.cpp:198, src/app.cpp:276,341)
    bool threw = false;
    try {
        recognizer.featureMatching();
    } catch (const char *s) {
        threw = std::string(s).find("No faces in database") != std::string::npos;
    }
    if (!threw) return 20;

    // a deterministic noise frame
    cv::Mat frame(H, W, CV_8UC3);
    unsigned s = 12345;
    for (size_t i = 0; i < static_cast<size_t>(H) * W * 3; ++i) {
        s = s * 1664525u + 1013904223u;
        frame.data[i] = static_cast<unsigned char>(s >> 24);
    }
    // enrolment path (src/app.cpp:86-93, src/db.cpp:326-339): resized 112x112 image -> preprocessFace -> doInference
    cv::Mat face, input;
    cv::resize(frame, face, cv::Size(recShape[1], recShape[2]));
    recognizer.preprocessFace(face, input);
    float emb1[512], emb2[512];
    recognizer.doInference((float *)input.ptr<float>(0), emb1);
    recognizer.doInference((float *)input.ptr<float>(0), emb2, 1);
    double n2 = 0, diff = 0;
    for (int i = 0; i < 512; ++i) {
        n2 += emb1[i] * emb1[i];
        diff += std::fabs(emb1[i] - emb2[i]);
    }
    if (std::fabs(n2 - 1.0) > 1e-3 || diff != 0) return 21;

    // detection + recognition path (src/app.cpp:304-310)
    outputBbox = detector.findFace(frame);
    if (outputBbox.empty()) return 22;
    std::vector<struct CroppedFace> cropped;
    getCroppedFaces(frame, outputBbox, recShape[2], recShape[1], cropped);  // src/app.cpp:170
    if (cropped.size() != outputBbox.size()) return 23;
    recognizer.forward(frame, outputBbox);
    if (recognizer.croppedFaces.size() != outputBbox.size() || recognizer.croppedFaces[0].face.rows != 112) return 24;

    // gallery: face 0's embedding as user "u0", the enrolment embedding as "u1" (db.getEmbeddings, src/db.cpp:316-346)
    recognizer.resetEmbeddings();
    recognizer.initKnownEmbeds(3);
    std::vector<float> e0(recognizer.embeddings(), recognizer.embeddings() + 512);
    recognizer.addEmbedding("u1", emb1);
    recognizer.addEmbedding("u0", e0);
    recognizer.addEmbedding("u0-dup", e0);  // duplicate row: the first maximum must win
    recognizer.initMatMul();
    if (ArcFaceIR50::classCount != 3) return 25;
    float *sims = recognizer.featureMatching();
    std::vector<std::string> names;
    std::vector<float> best;
    std::tie(names, best) = recognizer.getOutputs(sims);
    if (names.size() != outputBbox.size() || names[0] != "u0" || std::fabs(best[0] - 1.f) > 1e-4) return 26;
    std::vector<std::string> names2;
    std::vector<float> best2;
    std::tie(names2, best2) = recognizer.match();  // fused GPU top-1 agrees with the dense path
    for (size_t i = 0; i < names.size(); ++i)
        if (names2[i] != names[i] || best2[i] != best[i]) return 27;
    // /reload (src/app.cpp:354-365)
    recognizer.resetEmbeddings();
    recognizer.initKnownEmbeds(1);
    recognizer.addEmbedding("only", emb1);
    recognizer.initMatMul();
    std::tie(names, best) = recognizer.getOutputs(recognizer.featureMatching());
    if (names[0] != "only") return 28;
    // enrolment without a reload: append / remove on the resident gallery through the MatMul extension
    {
        MatMul mm;
        mm.init(e0.data(), 1, 512);
        mm.append(emb1, 1);
        float s2[2];
        int64_t r2[2];
        mm.search(emb1, 1, 2, s2, r2);
        if (r2[0] != 1 || std::fabs(s2[0] - 1.f) > 1e-4) return 29;
        if (mm.remove(0) != 1) return 30;  // the last row (emb1) moved into slot 0
        mm.search(emb1, 1, 1, s2, r2);
        if (r2[0] != 0) return 31;
    }
    std::printf("DROPIN OK: %zu faces, first box (%d,%d,%d,%d) score %.4f, top-1 %s %.4f\n", outputBbox.size(), outputBbox[0].x1, outputBbox[0].y1,
                outputBbox[0].x2, outputBbox[0].y2, outputBbox[0].score, names2[0].c_str(), best2[0]);
    return 0;
}
