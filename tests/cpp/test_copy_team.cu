// tests/cpp/test_copy_team.cu -- TEST: the copy-thread team of csrc/host_stage.h (the pageable-buffer stager)
// on the CPU alone: many back-to-back copies of random sizes and alignments, every byte checked, including the
// non-temporal-store path's unaligned heads and tails.  No CUDA call is made (CopyTeam is plain C++ threads).
#include <cstdio>
#include <random>
#include <vector>

#include "../../secure-computation-library_b200/csrc/host_stage.h"

int main() {
  std::mt19937_64 rng(12345);
  std::vector<unsigned char> a((40u << 20) + 4096), b((40u << 20) + 4096);
  for (auto& x : a) x = (unsigned char)rng();
  long checks = 0;
  for (int helpers : {0, 1, 3, 7}) {
    sclgpu::CopyTeam team(helpers);
    for (int it = 0; it < 600; ++it) {
      size_t n = (rng() % (it % 5 == 0 ? (40u << 20) : (3u << 20))) + 1;
      const size_t so = rng() % 4096, dof = rng() % 4096;
      if (so + n > a.size()) n = a.size() - so;
      if (dof + n > b.size()) n = b.size() - dof;
      std::memset(b.data() + dof, 0xA5, n);
      const unsigned char before = dof ? b[dof - 1] : 0, after = dof + n < b.size() ? b[dof + n] : 0;
      team.copy(b.data() + dof, a.data() + so, n);
      if (std::memcmp(b.data() + dof, a.data() + so, n) != 0 || (dof && b[dof - 1] != before) ||
          (dof + n < b.size() && b[dof + n] != after)) {
        std::printf("MISMATCH helpers=%d it=%d n=%zu\n", helpers, it, n);
        return 1;
      }
      ++checks;
    }
  }
  // stream_copy alone on tiny sizes
  for (size_t n = 0; n < 700; ++n)
    for (size_t off = 0; off < 17; ++off) {
      std::memset(b.data(), 0, 1024);
      sclgpu::stream_copy(b.data() + off, a.data() + 3, n);
      if (std::memcmp(b.data() + off, a.data() + 3, n) != 0 || b[off + n] != 0) {
        std::printf("MISMATCH stream_copy n=%zu off=%zu\n", n, off);
        return 1;
      }
      ++checks;
    }
  std::printf("COPY_TEAM_OK checks=%ld\n", checks);
  return 0;
}
