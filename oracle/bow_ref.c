/* oracle/bow_ref.c — CPU restatement of the reference's bag-of-words transform.  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may call it; nothing under gslam_b200/ links or loads it.
 *
 * Follows GSLAM/core/Vocabulary.h:
 *   - the tree walk of Vocabulary::transform(feature, word_id, weight, nid, levelsup)             :1692-1736
 *     (children of node p are the rows p*k+1 .. p*k+childNum of the node-descriptor matrix, :1714-1716; the FIRST child with the
 *     strictly smallest distance wins, :1720-1724; the walk ends at the first node without children, :1731; the word id IS the
 *     node id, :1734; nid is the node met at level L - levelsup, 0 when that level is <= 0, :1699-1700,1728-1729)
 *   - the distance DistanceFactory::hamming32                                                        :485-491
 *   - the accumulation of Vocabulary::transform(features, BowVector&, FeatureVector&, levelsup)      :1558-1622
 *     (TF / TF_IDF: float += per occurrence in feature order, addWeight :357-369; IDF / BINARY: first occurrence only,
 *     addIfNotExist :371-380; stopped words (weight <= 0) are skipped, :1585; feature vector: node -> feature indices in feature
 *     order, addFeature :410-425; then either the division by the number of words (TF / TF_IDF without normalisation, :1592-1598) or
 *     the L1 / L2 normalisation with a double norm accumulated in ascending word order, normalize :382-408)
 *   - which scoring types normalise: L1_NORM, CHI_SQUARE, KL, BHATTACHARYYA -> L1; L2_NORM -> L2; DOT_PRODUCT -> none   :667-684
 * Pinned against the reference itself (oracle/_ref: Vocabulary::create / load / transform compiled from the reference headers) by
 * tests/test_oracle_bow.py on trained and on synthetic vocabularies.
 *
 * One case the reference leaves undefined is DEFINED here (and in the CUDA path): a leaf met ABOVE level L - levelsup (an
 * unbalanced tree) leaves the reference's `nid` uninitialised (:1579,1728); we report the leaf itself.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { W_TF_IDF = 0, W_TF = 1, W_IDF = 2, W_BINARY = 3 };                                           /* Vocabulary.h:88-94 */
enum { S_L1 = 0, S_L2 = 1, S_CHI = 2, S_KL = 3, S_BHATT = 4, S_DOT = 5 };                           /* Vocabulary.h:97-105 */

static int hamming32(const uint8_t* a, const uint8_t* b) {                                          /* Vocabulary.h:485-491 */
  uint64_t x[4], y[4];
  memcpy(x, a, 32); memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
         __builtin_popcountll(x[3] ^ y[3]);
}

/* one descriptor down the tree */
void orc_bow_word(int k, int L, const uint32_t* child_num, const float* weight, const uint8_t* desc32, const uint8_t* feat, int levelsup,
                  int64_t* word, float* w, int64_t* node) {
  const int nid_level = L - levelsup;
  int64_t nid = -1;
  if (nid_level <= 0) nid = 0;
  int64_t cur = 0;
  int level = 0;
  do {
    ++level;
    float best_d = 3.402823466e+38f;
    int64_t best = cur;
    const int64_t first = cur * k + 1;
    for (int64_t id = first; id < first + (int64_t)child_num[cur]; ++id) {
      const float d = (float)hamming32(feat, desc32 + 32 * id);
      if (d < best_d) { best_d = d; best = id; }
    }
    cur = best;
    if (level == nid_level) nid = cur;
  } while (child_num[cur] != 0);
  if (nid < 0) nid = cur; /* (our definition, see the header) */
  *word = cur; *w = weight[cur]; *node = nid;
}

typedef struct { int64_t key; int idx; } KeyIdx;
static int cmp_keyidx(const void* a, const void* b) {
  const KeyIdx* x = (const KeyIdx*)a; const KeyIdx* y = (const KeyIdx*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx - y->idx;
}

/* features [n][32] -> BowVector (words ascending, values) + FeatureVector flattened (node ascending, feature index ascending).
 * Per-feature outputs (may be NULL): f_word, f_node.  Returns the number of words; *n_fv = entries of the feature vector. */
int orc_bow_transform(int k, int L, int weighting, int scoring, const uint32_t* child_num, const float* weight, const uint8_t* desc32,
                      const uint8_t* feats, int n, int levelsup, int64_t* words, float* values, int64_t* fv_node, int32_t* fv_feat,
                      int* n_fv, int64_t* f_word, int64_t* f_node) {
  KeyIdx* bw = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)(n > 0 ? n : 1));
  KeyIdx* bn = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)(n > 0 ? n : 1));
  float* fw = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int m = 0;
  for (int i = 0; i < n; ++i) {
    int64_t word, node; float w;
    orc_bow_word(k, L, child_num, weight, desc32, feats + 32 * (size_t)i, levelsup, &word, &w, &node);
    if (f_word) f_word[i] = word;
    if (f_node) f_node[i] = node;
    if (w > 0) { bw[m].key = word; bw[m].idx = i; bn[m].key = node; bn[m].idx = i; fw[i] = w; ++m; }
  }
  qsort(bw, (size_t)m, sizeof(KeyIdx), cmp_keyidx);  /* std::map order; equal words stay in feature order */
  qsort(bn, (size_t)m, sizeof(KeyIdx), cmp_keyidx);
  int nw = 0;
  const int tf = weighting == W_TF || weighting == W_TF_IDF;
  for (int a = 0; a < m;) {
    int b = a;
    float v = fw[bw[a].idx];
    for (b = a + 1; b < m && bw[b].key == bw[a].key; ++b)
      if (tf) v += fw[bw[b].idx];
    words[nw] = bw[a].key; values[nw] = v; ++nw;
    a = b;
  }
  const int must = scoring != S_DOT;
  if (tf && nw > 0 && !must) {
    const double nd = (double)nw;
    for (int a = 0; a < nw; ++a) values[a] = (float)(values[a] / nd);
  }
  if (must) {
    double norm = 0.0;
    if (scoring == S_L2) { for (int a = 0; a < nw; ++a) norm += values[a] * values[a]; norm = sqrt(norm); }  /* float product, as written */
    else for (int a = 0; a < nw; ++a) norm += fabs(values[a]);
    if (norm > 0.0) for (int a = 0; a < nw; ++a) values[a] = (float)(values[a] / norm);
  }
  for (int a = 0; a < m; ++a) { fv_node[a] = bn[a].key; fv_feat[a] = bn[a].idx; }
  *n_fv = m;
  free(bw); free(bn); free(fw);
  return nw;
}
