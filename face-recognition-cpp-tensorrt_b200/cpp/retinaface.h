// Drop-in replacement for the reference's RetinaFace class (/root/reference/src/retinaface.h:17-49, src/retinaface.cpp).
// Same constructor and findFace signature; `engineFile` is the packed weight file (tools/pack_weights.py).
#ifndef RETINAFACE_H
#define RETINAFACE_H

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include <cassert>

#include "common.h"

struct anchorBox {
    float cx;
    float cy;
    float sx;
    float sy;
};

class RetinaFace {
  public:
    RetinaFace(TRTLogger gLogger, const std::string engineFile, int frameWidth, int frameHeight, std::string inputName,
               std::vector<std::string> outputNames, std::vector<int> inputShape, int maxBatchSize, int maxFacesPerScene, float nms_threshold,
               float bbox_threshold) {
        (void)gLogger;
        (void)inputName;    // TensorRT binding names: accepted and ignored
        (void)outputNames;
        assert(inputShape.size() == 3);  // src/retinaface.cpp:8
        m_frameWidth = frameWidth;
        m_frameHeight = frameHeight;
        m_maxFacesPerScene = maxFacesPerScene;
        if (!fileExists(engineFile)) throw std::logic_error("Cant find engine file");  // src/retinaface.cpp:53
        std::cout << "[INFO] Loading RetinaFace Engine...\n";
        frCheck(fr_detector_create(engineFile.c_str(), inputShape[1], inputShape[2], frameHeight, frameWidth, maxBatchSize > 0 ? maxBatchSize : 1,
                                   maxFacesPerScene, nms_threshold, bbox_threshold, 0, 0, &m_detector));
        m_boxes.resize(maxFacesPerScene);
    }
    ~RetinaFace() { fr_detector_destroy(m_detector); }
    RetinaFace(const RetinaFace &) = delete;
    RetinaFace &operator=(const RetinaFace &) = delete;

    // preprocess + doInference + postprocessing (src/retinaface.cpp:147-152); returns a copy like the reference
    std::vector<struct Bbox> findFace(cv::Mat &img) {
        // The reference's preprocess resizes whatever it is given (src/retinaface.cpp:106-125) with the scales fixed at construction;
        // a frame of another size would be letterboxed wrongly there and read out of bounds here, so it is refused in every build
        // type (an assert would vanish under NDEBUG).
        if (img.empty() || img.rows != m_frameHeight || img.cols != m_frameWidth || img.type() != CV_8UC3)
            throw std::logic_error("RetinaFace::findFace: frame must be CV_8UC3 of the input_frameWidth x input_frameHeight given to the constructor");
        int count = 0;
        frCheck(fr_detector_run(m_detector, img.data, static_cast<int>(img.step), 1, reinterpret_cast<FrBbox *>(m_boxes.data()), &count, nullptr));
        m_outputBbox.assign(m_boxes.begin(), m_boxes.begin() + count);
        return m_outputBbox;
    }
    // batched extension (BASELINE.json north_star "RetinaFace::detect"): frames must be contiguous images of the frame size
    std::vector<std::vector<struct Bbox>> detect(const unsigned char *frames, int stride, int batch) {
        std::vector<Bbox> boxes(static_cast<size_t>(batch) * m_maxFacesPerScene);
        std::vector<int> counts(batch);
        frCheck(fr_detector_run(m_detector, frames, stride, batch, reinterpret_cast<FrBbox *>(boxes.data()), counts.data(), nullptr));
        std::vector<std::vector<struct Bbox>> out(batch);
        for (int b = 0; b < batch; ++b) out[b].assign(boxes.begin() + b * m_maxFacesPerScene, boxes.begin() + b * m_maxFacesPerScene + counts[b]);
        return out;
    }
    FrDetector *handle() const { return m_detector; }

  private:
    int m_frameWidth, m_frameHeight, m_maxFacesPerScene;
    FrDetector *m_detector = nullptr;
    std::vector<struct Bbox> m_boxes, m_outputBbox;
};

#endif  // RETINAFACE_H
