// Test harness (not product code): compiles the IoU routines of csrc/pp_boxes.cuh for the HOST, so that the float64
// evaluation the CUDA kernels run can be checked against the oracle on a machine without a GPU
// (tests/test_iou_math_host.py).  usage: iou_host {box3d|bev} in.bin out.bin
//   box3d: in = int32 n, m, then (n,8,3) and (m,8,3) float32 corners; out = (n,m) float32 volume, then (n,m) float32 iou
//   bev:   in = int32 n, m, then (n,9) and (m,9) float32 boxes;       out = (n,m) float32 iou
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#define PP_B3_FN __host__ __device__ inline
#include "pp_boxes.cuh"
using namespace pp;

static Box3 box_from_corners(const float *c)
{
    Box3 b;
    const int nb[3] = {1, 3, 4};
    for (int k = 0; k < 3; ++k) {
        b.o[k] = c[k];
        for (int j = 0; j < 3; ++j) b.e[j][k] = c[nb[j] * 3 + k] - c[k];
    }
    return b;
}
static RRect rect_from_box9(const float *b)
{
    RRect r;
    r.cx = b[0]; r.cy = b[1]; r.hx = 0.5f * b[3]; r.hy = 0.5f * b[4];
    r.s = sinf(b[8]); r.c = cosf(b[8]);
    return r;
}

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    const bool bev = !strcmp(argv[1], "bev");
    const int w = bev ? 9 : 24;
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 3;
    int n = 0, m = 0;
    if (fread(&n, 4, 1, f) != 1 || fread(&m, 4, 1, f) != 1) return 4;
    std::vector<float> a((size_t)n * w), b((size_t)m * w);
    if (fread(a.data(), 4, a.size(), f) != a.size() || fread(b.data(), 4, b.size(), f) != b.size()) return 4;
    fclose(f);
    std::vector<float> vol((size_t)n * m), iou((size_t)n * m);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            if (bev) {
                iou[(size_t)i * m + j] = rrect_iou(rect_from_box9(&a[(size_t)i * 9]), rect_from_box9(&b[(size_t)j * 9]));
            } else {
                float v;
                iou[(size_t)i * m + j] = box3_iou(box_from_corners(&a[(size_t)i * 24]), box_from_corners(&b[(size_t)j * 24]), &v);
                vol[(size_t)i * m + j] = v;
            }
        }
    f = fopen(argv[3], "wb");
    if (!f) return 3;
    if (!bev) fwrite(vol.data(), 4, vol.size(), f);
    fwrite(iou.data(), 4, iou.size(), f);
    fclose(f);
    return 0;
}
