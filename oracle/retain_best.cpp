// ORACLE (test infrastructure, not product code).
// Replays cv::KeyPointsFilter::retainBest's two std-algorithm calls on (response, original-index) pairs so the
// numpy oracle can reproduce cv2's keypoint ORDER: cv2 4.13 runs
//   std::nth_element(begin, begin+n-1, end, response-greater)
//   std::partition(begin+n, end, response >= boundary)
// with libstdc++ (g++ 13 here); the resulting order is implementation-defined, so the oracle calls the very same
// library algorithms.  Reference call sites whose output order depends on this: ORB::detectAndCompute via
// /root/reference/scripts/visual_odometry_v3.py:373.
#include <algorithm>
#include <cstdint>
#include <vector>

struct Item { float response; int32_t idx; };

extern "C" int oracle_retain_best(const float* response, int count, int n_points, int32_t* out_idx) {
    std::vector<Item> v(count);
    for (int i = 0; i < count; ++i) { v[i].response = response[i]; v[i].idx = i; }
    if (n_points >= 0 && count > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(),
                         [](const Item& a, const Item& b) { return a.response > b.response; });
        float boundary = v[n_points - 1].response;
        auto new_end = std::partition(v.begin() + n_points, v.end(),
                                      [boundary](const Item& a) { return a.response >= boundary; });
        v.resize(new_end - v.begin());
    }
    for (size_t i = 0; i < v.size(); ++i) out_idx[i] = v[i].idx;
    return (int)v.size();
}
