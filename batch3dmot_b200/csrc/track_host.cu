// Host-side (CPU) agglomerative track clustering for batched inference: create_trajectories(mode='hier')
// (reference batch_3dmot/predict.py:308-373) and the track-id numbering (:437-446) over the surviving edges of
// MANY scenes at once. The steps before it (window-score averaging, thresholds, best-in / best-out filter) are
// device tensor ops (batch3dmot_b200/tracking.py); what is left is inherently sequential — edges in stable
// descending-score order, each either opening a track, extending one at its head / tail, or joining two — so it
// runs here as one native loop over host arrays instead of a Python loop per scene. Scenes are disjoint node
// sets, so clustering the union in one global score order gives every scene exactly the tracks (and the track
// insertion order) that its own run would give: the per-scene edge sequence is a subsequence of the global one.
#include <algorithm>
#include <numeric>
#include <vector>

#include "b3d_common.cuh"

extern "C" int b3d_hier_tracks_host(const int64_t* e_out, const int64_t* e_in, const double* score, int64_t m,
                                    const int64_t* node_class, const int32_t* scene_of_node, int64_t n,
                                    int32_t n_scenes, const double* thresholds, int32_t n_classes,
                                    int64_t* track_id, int64_t* track_pos, int64_t* tracks_per_scene) {
  using namespace b3d;
  if (m < 0 || n < 0 || n_scenes < 1 || !node_class || !thresholds || !track_id || !track_pos || !tracks_per_scene ||
      (m > 0 && (!e_out || !e_in || !score)))
    return bad_arg("b3d_hier_tracks_host");
  std::vector<int64_t> order((size_t)m);
  std::iota(order.begin(), order.end(), (int64_t)0);
  // stable descending score == sorted(..., key=score, reverse=True) on a list in insertion order (predict.py:313)
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return score[a] > score[b]; });
  std::vector<int64_t> vis((size_t)n, -1), nxt((size_t)n, -1), head, tail;
  std::vector<char> alive;
  for (int64_t t : order) {
    const int64_t j = e_out[t], i = e_in[t];
    if (j < 0 || j >= n || i < 0 || i >= n) return bad_arg("b3d_hier_tracks_host: node id out of range");
    const int64_t cj = vis[j], ci = vis[i];
    if (cj < 0 && ci < 0) {                       // new track j -> i
      head.push_back(j); tail.push_back(i); alive.push_back(1);
      nxt[j] = i;
      vis[j] = vis[i] = (int64_t)head.size() - 1;
    } else if (cj < 0) {                          // prepend j to the track that starts with i
      if (head[ci] == i) { nxt[j] = i; head[ci] = j; vis[j] = ci; }
    } else if (ci < 0) {                          // append i to the track that ends with j
      if (tail[cj] == j) { nxt[j] = i; tail[cj] = i; vis[i] = cj; }
    } else {                                      // join: tail of one track to the head of another
      const int64_t c = node_class[i];
      if (c < 0 || c >= n_classes) return bad_arg("b3d_hier_tracks_host: class id out of range");
      if (tail[cj] == j && head[ci] == i && score[t] > thresholds[c]) {
        if (cj == ci) return bad_arg("b3d_hier_tracks_host: edge closes a cycle inside one track");
        nxt[j] = i;
        for (int64_t node = i; node >= 0; node = nxt[node]) vis[node] = cj;
        tail[cj] = tail[ci];
        alive[ci] = 0;
      }
    }
  }
  std::fill(track_id, track_id + n, (int64_t)-1);
  std::fill(track_pos, track_pos + n, (int64_t)-1);
  std::fill(tracks_per_scene, tracks_per_scene + n_scenes, (int64_t)0);
  for (size_t c = 0; c < head.size(); ++c) {
    if (!alive[c]) continue;
    const int32_t s = scene_of_node ? scene_of_node[head[c]] : 0;
    if (s < 0 || s >= n_scenes) return bad_arg("b3d_hier_tracks_host: scene id out of range");
    const int64_t tid = tracks_per_scene[s]++;    // position of the track among its scene's tracks (predict.py:438)
    int64_t pos = 0;
    for (int64_t node = head[c]; node >= 0; node = nxt[node]) { track_id[node] = tid; track_pos[node] = pos++; }
  }
  return 0;
}
