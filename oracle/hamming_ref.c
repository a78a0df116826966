/*
 * oracle/hamming_ref.c — CPU restatement of the 256-bit Hamming brute-force match.  TEST INFRASTRUCTURE ONLY:
 * imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product.
 *
 * Follows:
 *   distance  : GSLAM::Vocabulary::DistanceFactory::hamming32   GSLAM/core/Vocabulary.h:485-491
 *               (4 x uint64 XOR + popcount, returned as float 0..256; here as int).
 *   argmin    : cv::BFMatcher(NORM_HAMMING).match / knnMatch(k=2) semantics probed in SURVEY.md App. A.7:
 *               ties -> lowest train index; 2nd neighbour in (distance, index) order.
 * Pinned against: oracle/_ref (the reference's own hamming32 compiled from Vocabulary.h) and cv2.BFMatcher fixtures
 * in tests/golden (tests/test_oracle_hamming.py).
 */
#include <stdint.h>
#include <string.h>

int orc_hamming256(const uint8_t* a, const uint8_t* b) {
  uint64_t x[4], y[4];
  memcpy(x, a, 32);
  memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
         __builtin_popcountll(x[3] ^ y[3]);
}

/* best_idx/best_dist/second_dist: nq entries each (any may be NULL). */
void orc_match_hamming(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* best_idx,
                       int32_t* best_dist, int32_t* second_dist) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int i = 0; i < nq; ++i) {
    int b0 = -1, d0 = 257, d1 = 257;
    const uint8_t* q = query + (size_t)i * 32;
    for (int j = 0; j < nt; ++j) {
      int d = orc_hamming256(q, train + (size_t)j * 32);
      if (d < d0) { /* strictly better: previous best becomes 2nd (it has a lower index than any later equal one) */
        d1 = d0;
        d0 = d;
        b0 = j;
      } else if (d < d1) {
        d1 = d;
      }
    }
    if (best_idx) best_idx[i] = b0;
    if (best_dist) best_dist[i] = d0;
    if (second_dist) second_dist[i] = d1;
  }
}

/* Rectified-stereo row-band match (left = query, right = train): same distance, same (distance, index) order and tie rule as
 * above, restricted to the right keypoints with |y_R - y_L| <= band and min_disp <= x_L - x_R <= max_disp (float compares on the
 * level-0 pixel coordinates GSLAM::KeyPoint::pt, GSLAM/core/Map.h:180-194).  The reference has no stereo matcher (SURVEY.md 8f-1:
 * "stereo row-band Hamming match" is a next row); this is OUR definition -- parity unpinned by reference tests, pinned against an
 * independent numpy restatement in tests/test_oracle_hamming.py.  kps_*: records of 28 bytes, x at offset 0, y at offset 4. */
void orc_match_stereo(const uint8_t* kps_left, const uint8_t* desc_left, int nl, const uint8_t* kps_right, const uint8_t* desc_right, int nr,
                      float band, float min_disp, float max_disp, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
  for (int i = 0; i < nl; ++i) {
    float qx, qy;
    memcpy(&qx, kps_left + (size_t)i * 28, 4);
    memcpy(&qy, kps_left + (size_t)i * 28 + 4, 4);
    int b0 = -1, d0 = 257, d1 = 257;
    for (int j = 0; j < nr; ++j) {
      float px, py;
      memcpy(&px, kps_right + (size_t)j * 28, 4);
      memcpy(&py, kps_right + (size_t)j * 28 + 4, 4);
      const float dy = py - qy, disp = qx - px;
      if (!((dy < 0 ? -dy : dy) <= band && disp >= min_disp && disp <= max_disp)) continue;
      const int d = orc_hamming256(desc_left + (size_t)i * 32, desc_right + (size_t)j * 32);
      if (d < d0) { d1 = d0; d0 = d; b0 = j; }
      else if (d < d1) d1 = d;
    }
    if (best_idx) best_idx[i] = b0;
    if (best_dist) best_dist[i] = d0;
    if (second_dist) second_dist[i] = d1;
  }
}
