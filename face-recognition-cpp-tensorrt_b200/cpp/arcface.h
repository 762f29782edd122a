// Drop-in replacement for the reference's ArcFaceIR50 class, CroppedFace and getCroppedFaces
// (/root/reference/src/arcface.h:11-60, src/arcface.cpp). `engineFile` is the packed weight file (tools/pack_weights.py).
#ifndef ARCFACE_H
#define ARCFACE_H

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include <algorithm>
#include <cassert>
#include <cstring>
#include <tuple>

#include "common.h"
#include "matmul.h"

struct CroppedFace {
    cv::Mat face;     // 112x112 BGR u8 crop
    cv::Mat faceMat;  // working copy; after ArcFaceIR50::forward the reference leaves the normalised CHW float tensor here
    int x1, y1, x2, y2;
};

// crop + bicubic resize on the CPU through OpenCV, exactly the reference's statement sequence (src/arcface.cpp:3-17)
inline void getCroppedFaces(cv::Mat frame, std::vector<struct Bbox> &outputBbox, int resize_w, int resize_h,
                            std::vector<struct CroppedFace> &croppedFaces) {
    croppedFaces.clear();
    for (std::vector<struct Bbox>::iterator it = outputBbox.begin(); it != outputBbox.end(); it++) {
        cv::Rect facePos(cv::Point((*it).y1, (*it).x1), cv::Point((*it).y2, (*it).x2));
        cv::Mat tempCrop = frame(facePos);
        struct CroppedFace currFace;
        cv::resize(tempCrop, currFace.faceMat, cv::Size(resize_h, resize_w), 0, 0, cv::INTER_CUBIC);
        currFace.face = currFace.faceMat.clone();
        currFace.x1 = it->x1;
        currFace.y1 = it->y1;
        currFace.x2 = it->x2;
        currFace.y2 = it->y2;
        croppedFaces.push_back(currFace);
    }
}

// `static int classCount` (src/arcface.h:39, defined in src/arcface.cpp:19): header-only C++11 definition through a template base
template <class Tag>
struct ArcFaceStatics {
    static int classCount;
};
template <class Tag>
int ArcFaceStatics<Tag>::classCount = 0;

class ArcFaceIR50 : public ArcFaceStatics<void> {
  public:
    ArcFaceIR50(TRTLogger gLogger, const std::string engineFile, int frameWidth, int frameHeight, std::string inputName, std::string outputName,
                std::vector<int> inputShape, int outputDim, int maxBatchSize, int maxFacesPerScene, float knownPersonThreshold) {
        (void)gLogger;
        (void)inputName;
        (void)outputName;
        assert(inputShape.size() == 3);  // src/arcface.cpp:25
        m_frameWidth = frameWidth;
        m_frameHeight = frameHeight;
        m_INPUT_C = inputShape[0];
        m_INPUT_H = inputShape[1];
        m_INPUT_W = inputShape[2];
        m_OUTPUT_D = outputDim;
        m_maxBatchSize = maxBatchSize;
        m_maxFacesPerScene = maxFacesPerScene;
        m_knownPersonThresh = knownPersonThreshold;
        croppedFaces.reserve(maxFacesPerScene);
        m_embeds.resize(static_cast<size_t>(maxFacesPerScene) * m_OUTPUT_D);
        if (!fileExists(engineFile)) throw std::logic_error("Cant find engine file");  // src/arcface.cpp:67
        std::cout << "[INFO] Loading ArcFace Engine...\n";
        frCheck(fr_embedder_create(engineFile.c_str(), std::max(std::max(maxBatchSize, maxFacesPerScene), 1), 0, &m_embedder));
    }
    ~ArcFaceIR50() { fr_embedder_destroy(m_embedder); }
    ArcFaceIR50(const ArcFaceIR50 &) = delete;
    ArcFaceIR50 &operator=(const ArcFaceIR50 &) = delete;

    // BGR2RGB, float, (x - 127.5) * 0.0078125, planar CHW appended to `output` (src/arcface.cpp:105-114)
    void preprocessFace(cv::Mat &face, cv::Mat &output) {
        const int H = face.rows, W = face.cols;
        cv::Mat planes(3 * H, W, CV_32F);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < H; ++r) {
                const unsigned char *src = face.ptr<unsigned char>(r);
                float *dst = planes.ptr<float>(c * H + r);
                for (int x = 0; x < W; ++x) dst[x] = (static_cast<float>(src[x * 3 + (2 - c)]) - 127.5f) * 0.0078125f;
            }
        output.push_back(planes);
    }
    void doInference(float *input, float *output) { frCheck(fr_embedder_run(m_embedder, input, 1, output)); }
    void doInference(float *input, float *output, int batchSize) { frCheck(fr_embedder_run(m_embedder, input, batchSize, output)); }
    // alias named in BASELINE.json's north_star
    void extract(float *input, float *output, int batchSize) { doInference(input, output, batchSize); }

    void addEmbedding(const std::string className, float embedding[]) {
        classNames.push_back(className);
        std::copy(embedding, embedding + m_OUTPUT_D, m_knownEmbeds.begin() + static_cast<size_t>(classCount) * m_OUTPUT_D);
        classCount++;
    }
    void addEmbedding(const std::string className, std::vector<float> embedding) {
        classNames.push_back(className);
        std::copy(embedding.begin(), embedding.end(), m_knownEmbeds.begin() + static_cast<size_t>(classCount) * m_OUTPUT_D);
        classCount++;
    }
    // crop + resize + preprocess + embed on the GPU; row i of the embeddings = face i (src/arcface.cpp:166-187 without its
    // chunk-offset and m_embed overflow bugs). croppedFaces[i].face = the BGR crop, like the reference.
    void forward(cv::Mat frame, std::vector<struct Bbox> outputBbox) {
        croppedFaces.clear();
        const int n = static_cast<int>(outputBbox.size());
        if (n == 0) return;
        if (static_cast<size_t>(n) * m_OUTPUT_D > m_embeds.size()) m_embeds.resize(static_cast<size_t>(n) * m_OUTPUT_D);
        std::vector<unsigned char> crops(static_cast<size_t>(n) * m_INPUT_H * m_INPUT_W * 3);
        frCheck(fr_embedder_run_boxes(m_embedder, frame.data, frame.rows, frame.cols, static_cast<int>(frame.step),
                                      reinterpret_cast<const FrBbox *>(outputBbox.data()), n, m_embeds.data(), crops.data()));
        for (int i = 0; i < n; ++i) {
            CroppedFace f;
            f.face = cv::Mat(m_INPUT_H, m_INPUT_W, CV_8UC3);
            std::memcpy(f.face.data, crops.data() + static_cast<size_t>(i) * m_INPUT_H * m_INPUT_W * 3, static_cast<size_t>(m_INPUT_H) * m_INPUT_W * 3);
            f.faceMat = f.face;
            f.x1 = outputBbox[i].x1;
            f.y1 = outputBbox[i].y1;
            f.x2 = outputBbox[i].x2;
            f.y2 = outputBbox[i].y2;
            croppedFaces.push_back(f);
        }
    }
    // dense similarities, row-major faces x classCount, valid until the next call (the reference leaks a new[] per call)
    float *featureMatching() {
        if (classNames.size() > 0 && croppedFaces.size() > 0) {
            m_outputs.resize(croppedFaces.size() * static_cast<size_t>(classCount));
            matmul.calculate(m_embeds.data(), static_cast<int>(croppedFaces.size()), m_outputs.data());
        } else {
            throw "Feature matching: No faces in database or no faces found";  // src/arcface.cpp:198
        }
        return m_outputs.data();
    }
    // first maximum per face (src/arcface.cpp:203-217)
    std::tuple<std::vector<std::string>, std::vector<float>> getOutputs(float *output_sims) {
        std::vector<std::string> names;
        std::vector<float> sims;
        for (size_t i = 0; i < croppedFaces.size(); ++i) {
            const float *row = output_sims + i * classCount;
            const int argmax = static_cast<int>(std::max_element(row, row + classCount) - row);
            names.push_back(classNames[argmax]);
            sims.push_back(row[argmax]);
        }
        return std::make_tuple(names, sims);
    }
    // fused featureMatching + getOutputs on the GPU (no similarity matrix): the fast path
    std::tuple<std::vector<std::string>, std::vector<float>> match() {
        if (classNames.empty() || croppedFaces.empty()) throw "Feature matching: No faces in database or no faces found";
        const int n = static_cast<int>(croppedFaces.size());
        std::vector<float> sims(n);
        std::vector<int64_t> rows(n);
        matmul.search(m_embeds.data(), n, 1, sims.data(), rows.data());
        std::vector<std::string> names;
        for (int i = 0; i < n; ++i) names.push_back(classNames[rows[i]]);
        return std::make_tuple(names, sims);
    }
    void resetEmbeddings() {
        classCount = 0;
        classNames.clear();
    }
    void initKnownEmbeds(int num) { m_knownEmbeds.assign(static_cast<size_t>(num) * m_OUTPUT_D, 0.f); }
    void initMatMul() { matmul.init(m_knownEmbeds.data(), classCount, m_OUTPUT_D); }
    // debug drawing (src/arcface.cpp:219-231): rectangles and labels need OpenCV's drawing module; kept as a no-op hook
    void visualize(cv::Mat &image, std::vector<std::string> names, std::vector<float> sims) {
        (void)image;
        (void)names;
        (void)sims;
    }
    const float *embeddings() const { return m_embeds.data(); }
    FrEmbedder *handle() const { return m_embedder; }

    std::vector<struct CroppedFace> croppedFaces;

  private:
    int m_frameWidth, m_frameHeight, m_INPUT_C, m_INPUT_H, m_INPUT_W, m_OUTPUT_D, m_maxBatchSize, m_maxFacesPerScene;
    float m_knownPersonThresh;
    std::vector<float> m_embeds, m_knownEmbeds, m_outputs;
    std::vector<std::string> classNames;
    FrEmbedder *m_embedder = nullptr;
    MatMul matmul;
};

#endif  // ARCFACE_H
