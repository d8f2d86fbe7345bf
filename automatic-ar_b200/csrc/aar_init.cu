// aar_init.cu — host side of the C ABI in include/aar_init.h: the Initializer of the reference
// (/root/reference/libs/initializer.cpp) with its arithmetic on the device (aar_init.cuh) and its integer
// bookkeeping + spanning tree on the host.  No CPU fallback: every entry point fails with AAR_ERR_CUDA without a device.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/aar_cuda.h"
#include "../../include/aar_init.h"
#include "aar_init.cuh"

extern "C" void aar_internal_set_error(const char *msg);     // aar_cuda.cu: the thread-local message behind aar_last_error()

using namespace aar;

namespace {

void ierr(const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    aar_internal_set_error(buf);
}
#define ICU(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { ierr("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return AAR_ERR_CUDA; } \
    } while (0)

// 4x4 host arithmetic of the few chained transforms (initializer.cpp:290-314): cv::Mat operator* and cv::Mat::inv()
// in OpenCV's operation order (left-associated sums of products; LU with partial pivoting on [A | I])
struct H4 { double a[16]; };
H4 h4_eye() { H4 m; for (int i = 0; i < 16; i++) m.a[i] = (i % 5 == 0) ? 1.0 : 0.0; return m; }
H4 h4_mul(const H4 &A, const H4 &B) {
    H4 C;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = A.a[i * 4] * B.a[j];
            for (int k = 1; k < 4; k++) s = s + A.a[i * 4 + k] * B.a[k * 4 + j];
            C.a[i * 4 + j] = s;
        }
    return C;
}
H4 h4_inv(const H4 &in) {
    double A[16], b[16];
    std::memcpy(A, in.a, sizeof A);
    for (int i = 0; i < 16; i++) b[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 4; i++) {
        int k = i;
        for (int j = i + 1; j < 4; j++) if (std::fabs(A[j * 4 + i]) > std::fabs(A[k * 4 + i])) k = j;
        if (std::fabs(A[k * 4 + i]) < DBL_EPSILON * 100) { H4 z; std::memset(z.a, 0, sizeof z.a); return z; }
        if (k != i) { for (int j = i; j < 4; j++) std::swap(A[i * 4 + j], A[k * 4 + j]); for (int j = 0; j < 4; j++) std::swap(b[i * 4 + j], b[k * 4 + j]); }
        const double d = -1 / A[i * 4 + i];
        for (int j = i + 1; j < 4; j++) {
            const double alpha = A[j * 4 + i] * d;
            for (int kk = i + 1; kk < 4; kk++) A[j * 4 + kk] += alpha * A[i * 4 + kk];
            for (int kk = 0; kk < 4; kk++) b[j * 4 + kk] += alpha * b[i * 4 + kk];
        }
    }
    for (int i = 3; i >= 0; i--)
        for (int j = 0; j < 4; j++) {
            double s = b[i * 4 + j];
            for (int k = i + 1; k < 4; k++) s -= A[i * 4 + k] * b[k * 4 + j];
            b[i * 4 + j] = s / A[i * 4 + i];
        }
    H4 R; std::memcpy(R.a, b, sizeof b); return R;
}
H4 h4_from12(const double *p) { H4 m = h4_eye(); for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m.a[4 * i + j] = p[3 * i + j]; m.a[4 * i + 3] = p[9 + i]; } return m; }
void h4_to12(const H4 &m, double *p) { for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[3 * i + j] = m.a[4 * i + j]; p[9 + i] = m.a[4 * i + 3]; } }

typedef std::map<int, std::map<int, std::pair<H4, double>>> Best;

// Initializer::make_mst (initializer.cpp:237-288): the reference's Prim variant on edge weights
void make_mst(int start, const std::set<int> &ids, const Best &adj, std::map<int, std::set<int>> &children) {
    struct Node { double distance; int parent; };
    std::map<int, Node> outside;
    for (int id : ids) outside[id] = Node{id == start ? 0.0 : std::numeric_limits<double>::max(), -1};
    while (!outside.empty()) {
        auto mn = outside.begin();
        for (auto it = outside.begin(); it != outside.end(); ++it) if (it->second.distance < mn->second.distance) mn = it;
        const int a = mn->first;
        for (auto it = outside.begin(); it != outside.end(); ++it) {
            const int b = it->first;
            if (a == b) continue;
            auto row = adj.find(std::min(a, b));
            if (row == adj.end()) continue;
            auto col = row->second.find(std::max(a, b));
            if (col == row->second.end()) continue;
            const double error = col->second.second;
            if (error < it->second.distance) {
                it->second.distance = error;
                if (it->second.parent != -1) children[it->second.parent].erase(b);
                children[a].insert(b);
                it->second.parent = a;
            }
        }
        outside.erase(mn);
    }
}
// Initializer::find_transforms_to_root (initializer.cpp:290-314)
void transforms_to_root(int root, const std::map<int, std::set<int>> &children, const Best &best, std::map<int, H4> &out) {
    out[root] = h4_eye();
    std::queue<int> q; q.push(root);
    while (!q.empty()) {
        const int parent = q.front();
        auto ch = children.find(parent);
        if (ch != children.end())
            for (int child : ch->second) {
                if (child < parent) out[child] = best.at(child).at(parent).first;
                else out[child] = h4_inv(best.at(parent).at(child).first);
                if (parent != root) out[child] = h4_mul(out[parent], out[child]);
                q.push(child);
            }
        q.pop();
    }
}

struct Edge { int id1, id2; long long len; double weight; };

template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { if (p) cudaFree(p); p = nullptr; n = count; return count ? cudaMalloc((void **)&p, count * sizeof(T)) : cudaSuccess; }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t s) { cudaError_t e = alloc(v.size()); if (e != cudaSuccess || v.empty()) return e; return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s); }
};

} // namespace

struct aar_init {
    int device = 0; cudaStream_t stream = nullptr; bool own_stream = false;
    int num_cams = 0, num_frames = 0, min_detections = 2, consensus_max = 0;
    long long N = 0;
    double marker_size = 0, threshold = 2.0;
    std::vector<int> det_frame, det_cam, det_marker;
    std::vector<uint8_t> active, ncand;
    std::vector<int> frame_first;                 // detections of frame f: [frame_first[f], frame_first[f + 1]) (file order)
    std::vector<uint8_t> frame_kept;
    std::set<int> cam_ids, marker_ids;
    std::vector<int> marker_list;                 // ascending marker ids; index = rank
    std::map<int, int> marker_rank;
    int root_cam = -1, root_marker = -1;
    std::map<int, H4> to_root_cam, to_root_marker, object_T;
    std::vector<Edge> edges_cam, edges_marker;
    DevBuf<float> d_xy, d_err; DevBuf<int> d_cam, d_midx; DevBuf<uint8_t> d_active, d_ncand; DevBuf<double> d_K, d_dist, d_est;
    long long launches = 0; double ms[3] = {0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~aar_init() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); if (own_stream && stream) cudaStreamDestroy(stream); }
};

namespace {

// One batch of consensus lists: entries already on the device (tri), list boundaries on the host.  Returns per list the
// winner (position in the list, -1 if none), its error and its T.
int run_consensus(aar_init *h, const DevBuf<double> &tri, const std::vector<long long> &seg_begin, std::vector<long long> &best, std::vector<double> &weight, std::vector<double> &best_T) {
    const int nseg = (int)seg_begin.size() - 1;
    best.assign(nseg, -1); weight.assign(nseg, 0.0); best_T.assign((size_t)nseg * 12, 0.0);
    if (nseg <= 0) return AAR_OK;
    std::vector<ConsJob> jobs; std::vector<long long> job_begin(nseg + 1, 0);
    for (int s = 0; s < nseg; s++) {
        const long long m = seg_begin[s + 1] - seg_begin[s];
        if (m > (1LL << 31) - CS_THREADS) { ierr("consensus list of %lld candidates is too long", m); return AAR_ERR_UNSUPPORTED; }
        for (long long i = 0; i < m; i += CS_THREADS) jobs.push_back(ConsJob{s, (int)i});
        job_begin[s + 1] = (long long)jobs.size();
    }
    if (jobs.empty()) return AAR_OK;
    if (jobs.size() >= (1ull << 31)) { ierr("too many consensus jobs"); return AAR_ERR_UNSUPPORTED; }
    DevBuf<ConsJob> d_jobs; DevBuf<long long> d_seg, d_jb, d_pidx, d_best; DevBuf<double> d_pval, d_w, d_bt;
    ICU(d_jobs.upload(jobs, h->stream)); ICU(d_seg.upload(seg_begin, h->stream)); ICU(d_jb.upload(job_begin, h->stream));
    ICU(d_pval.alloc(jobs.size())); ICU(d_pidx.alloc(jobs.size())); ICU(d_best.alloc(nseg)); ICU(d_w.alloc(nseg)); ICU(d_bt.alloc((size_t)nseg * 12));
    k_consensus<<<(unsigned)jobs.size(), CS_THREADS, 0, h->stream>>>(d_jobs.p, d_seg.p, tri.p, h->marker_size / 2, d_pval.p, d_pidx.p);
    k_consensus_pick<<<(nseg + 127) / 128, 128, 0, h->stream>>>(nseg, d_jb.p, d_seg.p, d_pval.p, d_pidx.p, tri.p, d_best.p, d_w.p, d_bt.p);
    h->launches += 2;
    ICU(cudaGetLastError());
    ICU(cudaMemcpyAsync(best.data(), d_best.p, nseg * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    ICU(cudaMemcpyAsync(weight.data(), d_w.p, nseg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ICU(cudaMemcpyAsync(best_T.data(), d_bt.p, (size_t)nseg * 12 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ICU(cudaStreamSynchronize(h->stream));
    return AAR_OK;
}

// positions of a list of n candidates kept under consensus_max = k: floor(i * n / k), i = 0 .. k-1 (everything if n <= k or k = 0)
inline bool kept_position(long long pos, long long n, long long k) {
    if (k <= 0 || n <= k) return true;
    const long long i = (pos * k + n - 1) / n;            // smallest i with i * n / k >= pos
    return i < k && i * n / k == pos;
}

// fill_transformation_sets for one kind (cameras: groups are the markers of a frame, members the cameras that saw them;
// markers: groups are the cameras of a frame, members the markers they saw) — integer part, in the reference's list order.
// Pass 1 counts the candidates of every (id1, id2), pass 2 records the kept ones: src = (2 * det_a + cand_a, 2 * det_b + cand_b).
int rig_consensus(aar_init *h, bool cams, Best &best, std::vector<Edge> &edges) {
    typedef std::pair<int, int> Key;
    // member ids -> ranks (cameras: the id; markers: rank among the marker ids), pair (rank1, rank2) -> dense table index
    const int K = cams ? h->num_cams : (int)h->marker_list.size();
    if ((long long)K * K > (1LL << 26)) { ierr("too many %s (%d) for the pair tables of the rig initialisation", cams ? "cameras" : "markers", K); return AAR_ERR_UNSUPPORTED; }
    auto rank_of_member = [&](int id) { return cams ? id : (int)(std::lower_bound(h->marker_list.begin(), h->marker_list.end(), id) - h->marker_list.begin()); };
    // Two passes over the frames, both threaded over contiguous frame ranges: pass 0 counts the candidates of every (id1, id2) per
    // thread; the prefix sums over the threads give every thread the list position its range starts at, so that pass 1 can decide
    // which candidates survive `consensus_max` and collect them independently; concatenating the threads' lists in order gives the
    // list of the sequential scan (25.6 M detections: 8.0 s single-threaded).
    const size_t KK = (size_t)K * K;
    int T = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(1u, std::thread::hardware_concurrency()), (size_t)16, ((size_t)1 << 24) / std::max<size_t>(KK, 1), (size_t)std::max(1, h->num_frames / 64)}));
    std::vector<std::vector<long long>> count_t((size_t)T, std::vector<long long>(KK, 0));
    std::vector<std::vector<std::vector<int2>>> lists_t((size_t)T, std::vector<std::vector<int2>>(KK));
    std::vector<long long> count(KK, 0);
    struct Mem { int group, member_rank; long long det; };
    auto scan = [&](int t, int pass) {
        const int f0 = (int)((long long)h->num_frames * t / T), f1 = (int)((long long)h->num_frames * (t + 1) / T);
        std::vector<long long> &cnt = count_t[(size_t)t];      // pass 0: counts of this range; pass 1: running list positions (start = prefix over the threads)
        std::vector<Mem> members;
        for (int f = f0; f < f1; f++) {
            if (!h->frame_kept[f]) continue;
            members.clear();
            for (long long d = h->frame_first[f]; d < h->frame_first[f + 1]; d++)
                if (h->ncand[d]) members.push_back(cams ? Mem{h->det_marker[d], h->det_cam[d], d} : Mem{h->det_cam[d], rank_of_member(h->det_marker[d]), d});
            std::stable_sort(members.begin(), members.end(), [](const Mem &a, const Mem &b) { return a.group != b.group ? a.group < b.group : a.member_rank < b.member_rank; });
            size_t g0 = 0;
            while (g0 < members.size()) {
                size_t g1 = g0;
                while (g1 < members.size() && members[g1].group == members[g0].group) g1++;
                // members of the group [g0, g1) ascending by id; distinct ids = the reference's `objects` map
                if (members[g1 - 1].member_rank != members[g0].member_rank)
                    for (size_t a0 = g0; a0 < g1;) {
                        size_t a1 = a0; while (a1 < g1 && members[a1].member_rank == members[a0].member_rank) a1++;
                        const size_t row = (size_t)members[a0].member_rank * K;
                        for (size_t a = a0; a < a1; a++)
                            for (int i = 0; i < h->ncand[members[a].det]; i++)
                                for (size_t b = a1; b < g1; b++) {
                                    const size_t key = row + (size_t)members[b].member_rank;
                                    const int nb = h->ncand[members[b].det];
                                    if (pass == 0) cnt[key] += nb;
                                    else {
                                        long long &pos = cnt[key];
                                        const long long n = count[key];
                                        for (int j = 0; j < nb; j++, pos++)
                                            if (kept_position(pos, n, h->consensus_max))
                                                lists_t[(size_t)t][key].push_back(make_int2((int)(2 * members[a].det + i), (int)(2 * members[b].det + j)));
                                    }
                                }
                        a0 = a1;
                    }
                g0 = g1;
            }
        }
    };
    auto run_pass = [&](int pass) {
        if (T == 1) { scan(0, pass); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(scan, t, pass);
        for (auto &x : th) x.join();
    };
    run_pass(0);
    for (size_t k = 0; k < KK; k++) { long long run = 0; for (int t = 0; t < T; t++) { const long long c = count_t[(size_t)t][k]; count_t[(size_t)t][k] = run; run += c; } count[k] = run; }
    run_pass(1);
    std::vector<std::vector<int2>> lists(KK);
    for (size_t k = 0; k < KK; k++) {
        if (!count[k]) continue;
        size_t n = 0; for (int t = 0; t < T; t++) n += lists_t[(size_t)t][k].size();
        lists[k].reserve(n);
        for (int t = 0; t < T; t++) lists[k].insert(lists[k].end(), lists_t[(size_t)t][k].begin(), lists_t[(size_t)t][k].end());
    }
    { std::vector<std::vector<std::vector<int2>>>().swap(lists_t); }
    std::vector<int2> src; std::vector<long long> seg_begin(1, 0), seg_count; std::vector<Key> keys;
    for (int r1 = 0; r1 < K; r1++)                              // (id1, id2) ascending: the iteration order of the reference's nested maps
        for (int r2 = 0; r2 < K; r2++) {
            const size_t k = (size_t)r1 * K + r2;
            if (lists[k].empty()) continue;
            src.insert(src.end(), lists[k].begin(), lists[k].end()); seg_begin.push_back((long long)src.size()); seg_count.push_back(count[k]);
            keys.push_back(cams ? Key(r1, r2) : Key(h->marker_list[(size_t)r1], h->marker_list[(size_t)r2]));
        }
    if (src.empty()) return AAR_OK;
    DevBuf<int2> d_src; DevBuf<double> tri;
    ICU(d_src.upload(src, h->stream)); ICU(tri.alloc(src.size() * TRI_DOUBLES));
    k_build_pairs<<<(unsigned)((src.size() + 127) / 128), 128, 0, h->stream>>>((long long)src.size(), cams ? 1 : 0, d_src.p, h->d_est.p, tri.p);
    h->launches++;
    std::vector<long long> bi; std::vector<double> w, bt;
    int rc = run_consensus(h, tri, seg_begin, bi, w, bt);
    if (rc) return rc;
    for (size_t s = 0; s < keys.size(); s++) {
        if (bi[s] < 0) { ierr("consensus of (%d, %d) has no finite candidate", keys[s].first, keys[s].second); return AAR_ERR_NUMERIC; }
        best[keys[s].first][keys[s].second] = std::make_pair(h4_from12(&bt[12 * s]), w[s]);
        edges.push_back(Edge{keys[s].first, keys[s].second, seg_count[s], w[s]});
    }
    return AAR_OK;
}

int upload_rig_table(aar_init *h, int n, const std::map<int, H4> &T, const std::vector<int> &ids /* id of table row i, or row index == id when empty */, DevBuf<double> &tab) {
    std::vector<double> t12((size_t)n * 12, 0.0); std::vector<uint8_t> has(n, 0);
    for (int i = 0; i < n; i++) {
        auto it = T.find(ids.empty() ? i : ids[i]);
        if (it != T.end()) { has[i] = 1; h4_to12(it->second, &t12[12 * (size_t)i]); }
    }
    DevBuf<double> d_t; DevBuf<uint8_t> d_has;
    ICU(d_t.upload(t12, h->stream)); ICU(d_has.upload(has, h->stream)); ICU(tab.alloc((size_t)n * 24));
    k_rig_tables<<<(n + 63) / 64, 64, 0, h->stream>>>(n, d_t.p, d_has.p, tab.p);
    h->launches++;
    ICU(cudaStreamSynchronize(h->stream));               // d_t / d_has die with this scope
    return AAR_OK;
}

bool rigid_rows(const double *T, int n) { for (int i = 0; i < n; i++) { const double *m = T + 16 * (size_t)i; if (m[12] != 0 || m[13] != 0 || m[14] != 0 || m[15] != 1) return false; } return true; }

} // namespace

extern "C" {

int aar_init_create(const aar_init_desc *d, aar_init **out) {
    if (!d || !out) { ierr("null argument"); return AAR_ERR_INVALID; }
    *out = nullptr;
    if (d->num_cams < 1 || d->num_frames < 0 || d->num_detections < 0 || !d->cam_K || !d->cam_dist || !(d->marker_size > 0)) { ierr("bad initializer description"); return AAR_ERR_INVALID; }
    if (d->num_detections >= (1LL << 30)) { ierr("more than 2^30 detections"); return AAR_ERR_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ierr("no CUDA device: the B200 path has no CPU fallback"); return AAR_ERR_CUDA; }
    std::unique_ptr<aar_init> h(new aar_init);
    h->device = d->device; h->num_cams = d->num_cams; h->num_frames = d->num_frames; h->N = d->num_detections; h->marker_size = d->marker_size;
    h->threshold = d->threshold > 0 ? d->threshold : 2.0; h->min_detections = d->min_detections > 0 ? d->min_detections : 2; h->consensus_max = d->consensus_max;
    const long long N = h->N;
    h->det_frame.assign(d->det_frame, d->det_frame + N); h->det_cam.assign(d->det_cam, d->det_cam + N); h->det_marker.assign(d->det_marker, d->det_marker + N);
    // file order: frames ascending, cameras ascending within a frame
    h->frame_first.assign(h->num_frames + 1, 0);
    for (long long i = 0; i < N; i++) {
        const int f = h->det_frame[i], c = h->det_cam[i];
        if (f < 0 || f >= h->num_frames || c < 0 || c >= h->num_cams) { ierr("detection %lld: frame / camera out of range", i); return AAR_ERR_INVALID; }
        if (i > 0 && (f < h->det_frame[i - 1] || (f == h->det_frame[i - 1] && c < h->det_cam[i - 1]))) { ierr("detections must be in aruco.detections order (frame, camera)"); return AAR_ERR_INVALID; }
        h->frame_first[f + 1]++;
    }
    for (int f = 0; f < h->num_frames; f++) h->frame_first[f + 1] += h->frame_first[f];
    // initializer.cpp:373-380: frames with fewer than min_detections detections (excluded cameras not counted) are skipped
    h->active.assign(N, 0); h->frame_kept.assign(h->num_frames, 0);
    for (int f = 0; f < h->num_frames; f++) {
        int num = 0;
        for (long long i = h->frame_first[f]; i < h->frame_first[f + 1]; i++) if (!(d->excluded_cams && d->excluded_cams[h->det_cam[i]])) num++;
        if (!(num >= h->min_detections)) continue;
        h->frame_kept[f] = 1;
        for (long long i = h->frame_first[f]; i < h->frame_first[f + 1]; i++)
            if (!(d->excluded_cams && d->excluded_cams[h->det_cam[i]])) h->active[i] = 1;
    }
    {   // get_cam_ids / get_marker_ids: one pass over the active detections (a std::set insert per detection dominated cfg 4)
        std::vector<uint8_t> cam_seen((size_t)h->num_cams, 0);
        std::vector<int> mk; mk.reserve(1024);
        int last_marker = INT32_MIN;
        for (long long i = 0; i < N; i++)
            if (h->active[i]) { cam_seen[(size_t)h->det_cam[i]] = 1; const int m = h->det_marker[i]; if (m != last_marker) { mk.push_back(m); last_marker = m; if (mk.size() >= (1u << 20)) { std::sort(mk.begin(), mk.end()); mk.erase(std::unique(mk.begin(), mk.end()), mk.end()); } } }
        for (int c = 0; c < h->num_cams; c++) if (cam_seen[(size_t)c]) h->cam_ids.insert(c);
        h->marker_ids.insert(mk.begin(), mk.end());
    }
    h->marker_list.assign(h->marker_ids.begin(), h->marker_ids.end());
    for (size_t i = 0; i < h->marker_list.size(); i++) h->marker_rank[h->marker_list[i]] = (int)i;
    std::vector<int> midx(N, 0);
    for (long long i = 0; i < N; i++) if (h->active[i]) midx[i] = (int)(std::lower_bound(h->marker_list.begin(), h->marker_list.end(), h->det_marker[i]) - h->marker_list.begin());

    ICU(cudaSetDevice(h->device));
    if (d->stream) h->stream = (cudaStream_t)d->stream; else { ICU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    ICU(cudaEventCreate(&h->ev0)); ICU(cudaEventCreate(&h->ev1));
    ICU(h->d_xy.alloc((size_t)N * 8)); ICU(h->d_cam.upload(h->det_cam, h->stream)); ICU(h->d_midx.upload(midx, h->stream)); ICU(h->d_active.upload(h->active, h->stream));
    if (N) ICU(cudaMemcpyAsync(h->d_xy.p, d->det_xy, (size_t)N * 8 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    ICU(h->d_K.alloc((size_t)9 * h->num_cams)); ICU(h->d_dist.alloc((size_t)5 * h->num_cams));
    ICU(cudaMemcpyAsync(h->d_K.p, d->cam_K, sizeof(double) * 9 * h->num_cams, cudaMemcpyHostToDevice, h->stream));
    ICU(cudaMemcpyAsync(h->d_dist.p, d->cam_dist, sizeof(double) * 5 * h->num_cams, cudaMemcpyHostToDevice, h->stream));
    ICU(h->d_est.alloc((size_t)N * 24)); ICU(h->d_err.alloc((size_t)N * 2)); ICU(h->d_ncand.alloc(N));
    h->ncand.assign(N, 0);
    if (N) {
        ICU(cudaEventRecord(h->ev0, h->stream));
        k_ippe<<<(unsigned)((N + 127) / 128), 128, 0, h->stream>>>(N, h->d_xy.p, h->d_cam.p, h->d_active.p, h->d_K.p, h->d_dist.p, (float)h->marker_size, h->threshold,
                                                                    h->d_est.p, h->d_err.p, h->d_ncand.p);
        h->launches++;
        ICU(cudaEventRecord(h->ev1, h->stream));
        ICU(cudaGetLastError());
        ICU(cudaMemcpyAsync(h->ncand.data(), h->d_ncand.p, N, cudaMemcpyDeviceToHost, h->stream));
        ICU(cudaStreamSynchronize(h->stream));
        float ms = 0; ICU(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->ms[0] = ms;
    }
    *out = h.release();
    return AAR_OK;
}

void aar_init_destroy(aar_init *h) { if (h) { cudaSetDevice(h->device); delete h; } }

int aar_init_get_estimations(aar_init *h, double *T, double *err, uint8_t *ncand) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    ICU(cudaSetDevice(h->device));
    const long long N = h->N;
    if (T) {
        std::vector<double> e12((size_t)N * 24);
        if (N) ICU(cudaMemcpy(e12.data(), h->d_est.p, e12.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < 2 * N; i++) {
            if (!h->ncand[i / 2]) { std::memset(T + 16 * i, 0, 128); continue; }
            const H4 m = h4_from12(&e12[12 * (size_t)i]); std::memcpy(T + 16 * i, m.a, 128);
        }
    }
    if (err) {
        std::vector<float> e((size_t)N * 2);
        if (N) ICU(cudaMemcpy(e.data(), h->d_err.p, e.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < 2 * N; i++) err[i] = h->ncand[i / 2] ? (double)e[i] : 0.0;
    }
    if (ncand) std::memcpy(ncand, h->ncand.data(), N);
    return AAR_OK;
}

int aar_init_transforms(aar_init *h) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    if (h->cam_ids.empty() || h->marker_ids.empty()) { ierr("no frame with enough detections"); return AAR_ERR_INVALID; }
    ICU(cudaSetDevice(h->device));
    ICU(cudaEventRecord(h->ev0, h->stream));
    h->edges_cam.clear(); h->edges_marker.clear(); h->to_root_cam.clear(); h->to_root_marker.clear();
    Best best_cam, best_marker;
    int rc = rig_consensus(h, true, best_cam, h->edges_cam);
    if (rc) return rc;
    rc = rig_consensus(h, false, best_marker, h->edges_marker);
    if (rc) return rc;
    ICU(cudaEventRecord(h->ev1, h->stream)); ICU(cudaEventSynchronize(h->ev1));
    float ms = 0; ICU(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->ms[1] = ms;
    std::map<int, std::set<int>> cam_tree, marker_tree;
    h->root_cam = *h->cam_ids.begin();
    make_mst(h->root_cam, h->cam_ids, best_cam, cam_tree);
    transforms_to_root(h->root_cam, cam_tree, best_cam, h->to_root_cam);
    h->root_marker = *h->marker_ids.begin();
    make_mst(h->root_marker, h->marker_ids, best_marker, marker_tree);
    transforms_to_root(h->root_marker, marker_tree, best_marker, h->to_root_marker);
    return AAR_OK;
}

int aar_init_set_rig(aar_init *h, int32_t nc, const int32_t *cam_ids, const double *cam_T, int32_t nm, const int32_t *marker_ids, const double *marker_T) {
    if (!h || nc < 0 || nm < 0 || (nc && (!cam_ids || !cam_T)) || (nm && (!marker_ids || !marker_T))) { ierr("bad argument"); return AAR_ERR_INVALID; }
    if (!rigid_rows(cam_T, nc) || !rigid_rows(marker_T, nm)) { ierr("transforms must have last row [0 0 0 1]"); return AAR_ERR_UNSUPPORTED; }
    h->to_root_cam.clear(); h->to_root_marker.clear();
    for (int i = 0; i < nc; i++) { H4 m; std::memcpy(m.a, cam_T + 16 * (size_t)i, 128); h->to_root_cam[cam_ids[i]] = m; }
    for (int i = 0; i < nm; i++) { H4 m; std::memcpy(m.a, marker_T + 16 * (size_t)i, 128); h->to_root_marker[marker_ids[i]] = m; }
    h->root_cam = nc ? h->to_root_cam.begin()->first : -1; h->root_marker = nm ? h->to_root_marker.begin()->first : -1;
    return AAR_OK;
}

int aar_init_object_transforms(aar_init *h) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    ICU(cudaSetDevice(h->device));
    h->object_T.clear();
    // candidate list of a frame: markers ascending, cameras ascending, detection order, candidates (initializer.cpp:76-110)
    // threaded over contiguous frame ranges; the threads' pieces are concatenated in order
    const int T = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(1u, std::thread::hardware_concurrency()), (size_t)16, (size_t)std::max(1, h->num_frames / 256)}));
    std::vector<std::vector<int>> src_t((size_t)T), len_t((size_t)T), frame_t((size_t)T);
    auto scan = [&](int t) {
        const int f0 = (int)((long long)h->num_frames * t / T), f1 = (int)((long long)h->num_frames * (t + 1) / T);
        std::vector<std::pair<std::pair<int, int>, long long>> members;
        std::vector<int> one;
        for (int f = f0; f < f1; f++) {
            if (!h->frame_kept[f]) continue;
            members.clear();
            for (long long d = h->frame_first[f]; d < h->frame_first[f + 1]; d++) if (h->ncand[d]) members.push_back({{h->det_marker[d], h->det_cam[d]}, d});
            std::stable_sort(members.begin(), members.end(), [](const std::pair<std::pair<int, int>, long long> &a, const std::pair<std::pair<int, int>, long long> &b) { return a.first < b.first; });
            one.clear();
            for (auto &m : members) for (int k = 0; k < h->ncand[m.second]; k++) one.push_back((int)(2 * m.second + k));
            const long long n = (long long)one.size();
            int kept = 0;
            for (long long pos = 0; pos < n; pos++) if (kept_position(pos, n, h->consensus_max)) { src_t[(size_t)t].push_back(one[pos]); kept++; }
            len_t[(size_t)t].push_back(kept); frame_t[(size_t)t].push_back(f);
        }
    };
    if (T == 1) scan(0);
    else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(scan, t); for (auto &x : th) x.join(); }
    std::vector<int> src; std::vector<long long> seg_begin(1, 0); std::vector<int> seg_frame;
    { size_t n = 0; for (auto &v : src_t) n += v.size(); src.reserve(n); }
    for (int t = 0; t < T; t++) {
        src.insert(src.end(), src_t[(size_t)t].begin(), src_t[(size_t)t].end());
        for (size_t i = 0; i < len_t[(size_t)t].size(); i++) { seg_begin.push_back(seg_begin.back() + len_t[(size_t)t][i]); seg_frame.push_back(frame_t[(size_t)t][i]); }
    }
    if (src.empty()) return AAR_OK;
    ICU(cudaEventRecord(h->ev0, h->stream));
    DevBuf<double> cam_tab, mk_tab, tri; DevBuf<int> d_src;
    int rc = upload_rig_table(h, h->num_cams, h->to_root_cam, std::vector<int>(), cam_tab);
    if (rc) return rc;
    rc = upload_rig_table(h, (int)h->marker_list.size(), h->to_root_marker, h->marker_list, mk_tab);
    if (rc) return rc;
    ICU(d_src.upload(src, h->stream)); ICU(tri.alloc(src.size() * TRI_DOUBLES));
    k_build_object<<<(unsigned)((src.size() + 127) / 128), 128, 0, h->stream>>>((long long)src.size(), d_src.p, h->d_est.p, h->d_cam.p, h->d_midx.p, cam_tab.p, mk_tab.p, tri.p);
    h->launches++;
    std::vector<long long> bi; std::vector<double> w, bt;
    rc = run_consensus(h, tri, seg_begin, bi, w, bt);
    if (rc) return rc;
    ICU(cudaEventRecord(h->ev1, h->stream)); ICU(cudaEventSynchronize(h->ev1));
    float ms = 0; ICU(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->ms[2] = ms;
    for (size_t s = 0; s < seg_frame.size(); s++) if (bi[s] >= 0) h->object_T[seg_frame[s]] = h4_from12(&bt[12 * s]);
    return AAR_OK;
}

int aar_init_counts(const aar_init *h, int32_t *c) {
    if (!h || !c) { ierr("null argument"); return AAR_ERR_INVALID; }
    c[0] = (int)h->cam_ids.size(); c[1] = (int)h->marker_ids.size(); c[2] = (int)h->to_root_cam.size(); c[3] = (int)h->to_root_marker.size();
    c[4] = (int)h->object_T.size(); c[5] = h->root_cam; c[6] = h->root_marker;
    return AAR_OK;
}
int aar_init_get_ids(const aar_init *h, int32_t *cam_ids, int32_t *marker_ids) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    if (cam_ids) std::copy(h->cam_ids.begin(), h->cam_ids.end(), cam_ids);
    if (marker_ids) std::copy(h->marker_ids.begin(), h->marker_ids.end(), marker_ids);
    return AAR_OK;
}
static void copy_out(const std::map<int, H4> &m, int32_t *ids, double *T) {
    size_t i = 0;
    for (auto &kv : m) { if (ids) ids[i] = kv.first; if (T) std::memcpy(T + 16 * i, kv.second.a, 128); i++; }
}
int aar_init_get_rig(const aar_init *h, int32_t *cam_ids, double *cam_T, int32_t *marker_ids, double *marker_T) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    copy_out(h->to_root_cam, cam_ids, cam_T); copy_out(h->to_root_marker, marker_ids, marker_T);
    return AAR_OK;
}
int aar_init_get_object_transforms(const aar_init *h, int32_t *frame_ids, double *T) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    copy_out(h->object_T, frame_ids, T);
    return AAR_OK;
}
int aar_init_edges(const aar_init *h, int32_t cams, int32_t capacity, int32_t *id1, int32_t *id2, int64_t *list_len, double *weight, int32_t *num_edges) {
    if (!h || !num_edges) { ierr("null argument"); return AAR_ERR_INVALID; }
    const std::vector<Edge> &e = cams ? h->edges_cam : h->edges_marker;
    *num_edges = (int)e.size();
    for (int i = 0; i < (int)e.size() && i < capacity; i++) {
        if (id1) id1[i] = e[i].id1; if (id2) id2[i] = e[i].id2; if (list_len) list_len[i] = e[i].len; if (weight) weight[i] = e[i].weight;
    }
    return AAR_OK;
}

int aar_init_consensus(int32_t device, double marker_size, int64_t n, const double *T, const double *T1inv, const double *T2inv, int32_t *best, double *weight) {
    if (n < 0 || (n && (!T || !T1inv || !T2inv)) || !best || !weight) { ierr("bad argument"); return AAR_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ierr("no CUDA device: the B200 path has no CPU fallback"); return AAR_ERR_CUDA; }
    *best = -1; *weight = 0;
    if (n == 0) return AAR_OK;
    if (!rigid_rows(T, (int)n) || !rigid_rows(T1inv, (int)n) || !rigid_rows(T2inv, (int)n)) { ierr("transforms must have last row [0 0 0 1]"); return AAR_ERR_UNSUPPORTED; }
    aar_init h; h.device = device; h.marker_size = marker_size;
    ICU(cudaSetDevice(device));
    ICU(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking)); h.own_stream = true;
    std::vector<double> tri((size_t)n * TRI_DOUBLES);
    for (int64_t i = 0; i < n; i++) {
        H4 a, b, c; std::memcpy(a.a, T + 16 * i, 128); std::memcpy(b.a, T1inv + 16 * i, 128); std::memcpy(c.a, T2inv + 16 * i, 128);
        h4_to12(a, &tri[(size_t)i * TRI_DOUBLES]); h4_to12(b, &tri[(size_t)i * TRI_DOUBLES + 12]); h4_to12(c, &tri[(size_t)i * TRI_DOUBLES + 24]);
    }
    DevBuf<double> d_tri; ICU(d_tri.upload(tri, h.stream));
    std::vector<long long> seg = {0, (long long)n}, bi; std::vector<double> w, bt;
    int rc = run_consensus(&h, d_tri, seg, bi, w, bt);
    if (rc) return rc;
    *best = (int32_t)bi[0]; *weight = w[0];
    return AAR_OK;
}

int aar_init_timings(const aar_init *h, double *ms, int64_t *launches) {
    if (!h) { ierr("null handle"); return AAR_ERR_INVALID; }
    if (ms) for (int i = 0; i < 3; i++) ms[i] = h->ms[i];
    if (launches) *launches = h->launches;
    return AAR_OK;
}

} // extern "C"
