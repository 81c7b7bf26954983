// hd_infer — native inference runner over libhd_b200.so.
//
// The role of the reference's C++ demo (PytorchToCpp/main.cpp:35-118: load a TorchScript module, normalise one image,
// forward twice and time the second pass, print "index / xmin / ymin / xmax / ymax / class / score" per detection),
// without LibTorch or OpenCV: the network runs through the C ABI (include/hd_b200.h) - hd_normalize_u8,
// hd_net_forward (eval mode: BatchNorm folded into the conv epilogues), hd_decode_nms - and the whole pass is also
// recorded into a CUDA graph and replayed (--iters) the way a serving loop would run it.
//
//   hd_infer -m model.hdw -i image.ppm [--topk 100] [--conf-th 0.2] [--nms-th 0.2] [--iters 100] [--csv out.csv]
//   hd_infer -m model.hdw --random 512      (synthetic image, seeded)
//
// model.hdw: written by real_time_helmet_detection_b200.export.export_weights (format documented there).
// image: binary PPM (P6, maxval 255), height and width multiples of 64 (the reference resizes to 512x512 with OpenCV
// first; resizing is left to the caller here). Normalisation: ImageNet mean / std as PytorchToCpp/main.cpp:10-11.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "hd_b200.h"

#define CK(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            std::exit(2);                                                                             \
        }                                                                                             \
    } while (0)
#define HD(expr)                                                                                  \
    do {                                                                                          \
        int rc_ = (expr);                                                                         \
        if (rc_ != 0) {                                                                           \
            std::fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #expr, rc_, hd_last_error()); \
            std::exit(3);                                                                         \
        }                                                                                         \
    } while (0)

static float* upload(std::ifstream& f, size_t n) {
    std::vector<float> h(n);
    f.read(reinterpret_cast<char*>(h.data()), static_cast<std::streamsize>(n * sizeof(float)));
    if (!f) { std::fprintf(stderr, "model file truncated\n"); std::exit(1); }
    float* d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(float)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

static bool load_ppm(const std::string& path, std::vector<uint8_t>& rgb, int& W, int& H) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P6") return false;
    int vals[3], got = 0;
    while (got < 3) {
        f >> std::ws;
        if (f.peek() == '#') { std::string line; std::getline(f, line); continue; }
        if (!(f >> vals[got])) return false;
        ++got;
    }
    f.get();  // the single whitespace byte after maxval
    if (vals[2] != 255) return false;
    W = vals[0]; H = vals[1];
    rgb.resize(static_cast<size_t>(W) * H * 3);
    f.read(reinterpret_cast<char*>(rgb.data()), static_cast<std::streamsize>(rgb.size()));
    return static_cast<bool>(f);
}

int main(int argc, char** argv) {
    std::string model, image, csv;
    int topk = 100, iters = 0, random_size = 0;
    float conf_th = 0.2f, nms_th = 0.2f;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(1); }
            return argv[++i];
        };
        if (a == "-m" || a == "--model") model = next();
        else if (a == "-i" || a == "--image") image = next();
        else if (a == "--random") random_size = std::atoi(next());
        else if (a == "--topk") topk = std::atoi(next());
        else if (a == "--conf-th") conf_th = static_cast<float>(std::atof(next()));
        else if (a == "--nms-th") nms_th = static_cast<float>(std::atof(next()));
        else if (a == "--iters") iters = std::atoi(next());
        else if (a == "--csv") csv = next();
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 1; }
    }
    if (model.empty() || (image.empty() && random_size <= 0)) {
        std::fprintf(stderr, "usage: hd_infer -m model.hdw (-i image.ppm | --random SIZE) [--topk K] [--conf-th T] "
                             "[--nms-th T] [--iters N] [--csv FILE]\n");
        return 1;
    }

    // ---- image
    std::vector<uint8_t> rgb;
    int W = 0, H = 0;
    if (!image.empty()) {
        if (!load_ppm(image, rgb, W, H)) { std::fprintf(stderr, "cannot read %s as a binary PPM (P6, maxval 255)\n", image.c_str()); return 1; }
    } else {
        W = H = random_size;
        rgb.resize(static_cast<size_t>(W) * H * 3);
        uint32_t s = 12345u;
        for (auto& v : rgb) { s = s * 1664525u + 1013904223u; v = static_cast<uint8_t>(s >> 24); }
    }
    std::printf("image.size: [%d x %d]\n", W, H);
    if (W % 64 || H % 64 || W <= 0 || H <= 0) { std::fprintf(stderr, "image width and height must be multiples of 64\n"); return 1; }

    // ---- model
    std::ifstream f(model, std::ios::binary);
    char magic[4];
    int32_t hdr[4];
    if (!f || !f.read(magic, 4) || std::memcmp(magic, "HDW1", 4) != 0 || !f.read(reinterpret_cast<char*>(hdr), 16)) {
        std::fprintf(stderr, "%s is not an HDW1 weight file\n", model.c_str());
        return 1;
    }
    const int S = hdr[0], in_ch = hdr[1], out_ch = hdr[2], n_units = hdr[3], C = out_ch - 4;
    hd_net* net = nullptr;
    HD(hd_net_create(S, in_ch, out_ch, &net));
    if (hd_net_num_units(net) != n_units) { std::fprintf(stderr, "model has %d units, the executor expects %d\n", n_units, hd_net_num_units(net)); return 1; }
    std::vector<hd_unit_ptrs> units(static_cast<size_t>(n_units));
    long long* nbt = nullptr;                   // num_batches_tracked is only written in training mode
    CK(cudaMalloc(&nbt, sizeof(long long)));
    CK(cudaMemset(nbt, 0, sizeof(long long)));
    for (auto& u : units) {
        int32_t d[5];
        f.read(reinterpret_cast<char*>(d), 20);
        std::memset(&u, 0, sizeof(u));
        u.w = upload(f, static_cast<size_t>(d[0]) * d[1] * d[2] * d[2]);
        if (d[3]) u.b = upload(f, d[0]);
        if (d[4]) {
            u.gamma = upload(f, d[0]); u.beta = upload(f, d[0]);
            u.running_mean = upload(f, d[0]); u.running_var = upload(f, d[0]);
            u.num_batches_tracked = nbt;
        }
    }
    hd_net_set_static_weights(net, 1);      // deployment: weights are packed to bf16 once, on the first forward
    std::printf("model load!\n");

    // ---- buffers
    const int h = H / 4, w = W / 4, n_out = S * topk;
    cudaStream_t stream;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    uint8_t* d_img = nullptr; float *d_x = nullptr, *d_logits = nullptr, *d_boxes = nullptr, *d_scores = nullptr;
    long long* d_cls = nullptr; int* d_count = nullptr; void *d_ws = nullptr, *d_scratch = nullptr;
    const size_t ws_bytes = hd_net_workspace_bytes(net, 1, H, W, 0);
    const size_t logits_n = static_cast<size_t>(S) * out_ch * h * w;
    CK(cudaMalloc(&d_img, rgb.size()));
    CK(cudaMalloc(&d_x, rgb.size() * sizeof(float)));
    CK(cudaMalloc(&d_logits, logits_n * sizeof(float)));
    CK(cudaMalloc(&d_ws, ws_bytes));
    CK(cudaMalloc(&d_scratch, hd_decode_scratch_bytes(1, S, C, h, w)));
    HD(hd_decode_scratch_init(d_scratch, 1, S, stream));       // once: hd_decode_nms leaves the counters zeroed
    CK(cudaMalloc(&d_boxes, static_cast<size_t>(n_out) * 4 * sizeof(float)));
    CK(cudaMalloc(&d_cls, static_cast<size_t>(n_out) * sizeof(long long)));
    CK(cudaMalloc(&d_scores, static_cast<size_t>(n_out) * sizeof(float)));
    CK(cudaMalloc(&d_count, sizeof(int)));
    uint8_t* h_img = nullptr;                                  // pinned: the H2D copy is part of the recorded graph
    CK(cudaHostAlloc(reinterpret_cast<void**>(&h_img), rgb.size(), cudaHostAllocDefault));
    std::memcpy(h_img, rgb.data(), rgb.size());
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    const long long bs = static_cast<long long>(logits_n), ss = static_cast<long long>(out_ch) * h * w;

    auto enqueue = [&]() {      // one prediction: H2D, normalise, network, decode + NMS - enqueue only
        CK(cudaMemcpyAsync(d_img, h_img, rgb.size(), cudaMemcpyHostToDevice, stream));
        HD(hd_normalize_u8(d_img, d_x, 1, H, W, mean, stdv, stream));
        HD(hd_net_forward(net, units.data(), n_units, d_x, d_logits, d_ws, ws_bytes, 1, H, W, 0, stream));
        HD(hd_decode_nms(d_logits, bs, ss, d_logits + static_cast<size_t>(C) * h * w, bs, ss,
                         d_logits + static_cast<size_t>(C + 2) * h * w, bs, ss, 1, S, C, h, w, topk, 4.0f, conf_th, nms_th, 0, 1,
                         1, d_scratch, d_boxes, d_cls, d_scores, d_count, stream));
    };
    // The first forwarding takes longer (module load, stream / event creation): forward twice, time the second
    // (PytorchToCpp/main.cpp:58-64).
    enqueue();
    CK(cudaStreamSynchronize(stream));
    auto t0 = std::chrono::high_resolution_clock::now();
    enqueue();
    CK(cudaStreamSynchronize(stream));
    auto t1 = std::chrono::high_resolution_clock::now();
    std::printf("Inference Time: %g(ms)\n", std::chrono::duration<double>(t1 - t0).count() * 1000);

    if (iters > 0) {            // serving loop: the same pass recorded once into a CUDA graph and replayed
        cudaGraph_t graph; cudaGraphExec_t exec;
        CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        enqueue();
        CK(cudaStreamEndCapture(stream, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphLaunch(exec, stream));
        CK(cudaStreamSynchronize(stream));
        auto g0 = std::chrono::high_resolution_clock::now();
        for (int i = 0; i < iters; ++i) {
            CK(cudaGraphLaunch(exec, stream));
            CK(cudaStreamSynchronize(stream));
        }
        auto g1 = std::chrono::high_resolution_clock::now();
        const double ms = std::chrono::duration<double>(g1 - g0).count() * 1000 / iters;
        std::printf("CUDA graph replay: %g(ms) per image, %g FPS (%d iterations, one sync each)\n", ms, 1000.0 / ms, iters);
        CK(cudaGraphExecDestroy(exec));
        CK(cudaGraphDestroy(graph));
    }

    // ---- results
    int count = 0;
    CK(cudaMemcpy(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> boxes(static_cast<size_t>(count) * 4), scores(count);
    std::vector<long long> cls(count);
    if (count > 0) {
        CK(cudaMemcpy(boxes.data(), d_boxes, boxes.size() * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cls.data(), d_cls, cls.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(scores.data(), d_scores, scores.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    std::printf("Prediction Result(box, class, score)\n");
    for (int i = 0; i < count; ++i)     // integer corners like the cv::Point of main.cpp:88-96
        std::printf("index: %d, xmin: %d, ymin: %d, xmax: %d, ymax: %d, class: %lld, score: %g\n", i,
                    static_cast<int>(boxes[4 * i]), static_cast<int>(boxes[4 * i + 1]), static_cast<int>(boxes[4 * i + 2]),
                    static_cast<int>(boxes[4 * i + 3]), cls[i], scores[i]);
    if (!csv.empty()) {
        FILE* o = std::fopen(csv.c_str(), "w");
        if (!o) { std::fprintf(stderr, "cannot write %s\n", csv.c_str()); return 1; }
        for (int i = 0; i < count; ++i)
            std::fprintf(o, "%.9g,%.9g,%.9g,%.9g,%lld,%.9g\n", boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3],
                         cls[i], scores[i]);
        std::fclose(o);
    }
    hd_net_destroy(net);
    return 0;
}
