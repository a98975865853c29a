// Note decoding on the host: frames -> [[onset_s, offset_s, midi], ...].
// Semantics of frame2note (MIR_ST500/utils.py:82-149): a sequential state machine over frames with
//   * onset  : p_on[i] >= thr (fp32 compare, the recipe passes 0-d fp32 tensors) and p_on[i] equal to the
//              maximum of the window [max(i-3,0), min(i+4, n-1)) -- the upper clamp to n-1 is the
//              reference's, so the last frame is never inside a window,
//   * offset : p_off[i] >= thr closes the open note,
//   * pitch  : mode of oct*12+pc over the note's frames (frames with oct==4 or pc==12 are skipped); ties are
//              resolved like CPython's max(set(c), key=c.count): the first maximal element in the iteration
//              order of a CPython set built by inserting the values in sequence.
// The set order is obtained by replaying CPython 3.12's open-addressing table (hash(n) = n for small ints).
#include <cstdint>
#include <vector>

#include "../../include/svt_b200.h"
#include "common.cuh"

namespace {

// Insertion-ordered replay of CPython's set table for non-negative small ints.
class SmallIntSetOrder {
 public:
  SmallIntSetOrder() : slots_(8, kEmpty), mask_(7), fill_(0) {}
  void add(int key) {
    size_t i = probe(slots_, mask_, key, /*stop_on_equal=*/true);
    if (slots_[i] == key) return;
    slots_[i] = key;
    if (static_cast<size_t>(++fill_) * 5 >= mask_ * 3) grow(static_cast<size_t>(fill_) * 4);
  }
  // slot order == iteration order of the Python set
  const std::vector<int>& slots() const { return slots_; }
  static constexpr int kEmpty = -1;

 private:
  static size_t probe(const std::vector<int>& tab, size_t mask, int key, bool stop_on_equal) {
    size_t perturb = static_cast<size_t>(key);
    size_t i = static_cast<size_t>(key) & mask;
    while (true) {
      const size_t limit = (i + 9 <= mask) ? 9 : 0;  // LINEAR_PROBES
      for (size_t j = 0; j <= limit; ++j) {
        const int cur = tab[i + j];
        if (cur == kEmpty || (stop_on_equal && cur == key)) return i + j;
      }
      perturb >>= 5;  // PERTURB_SHIFT
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  void grow(size_t min_used) {
    size_t size = 8;
    while (size <= min_used) size <<= 1;
    std::vector<int> fresh(size, kEmpty);
    for (int key : slots_)
      if (key != kEmpty) fresh[probe(fresh, size - 1, key, false)] = key;
    slots_.swap(fresh);
    mask_ = size - 1;
  }
  std::vector<int> slots_;
  size_t mask_;
  int fill_;
};

int pitch_mode(const std::vector<int>& pitches) {
  SmallIntSetOrder order;
  int counts[64] = {0};
  for (int p : pitches) {
    order.add(p);
    ++counts[p];
  }
  int best = -1, best_count = 0;
  for (int key : order.slots())
    if (key != SmallIntSetOrder::kEmpty && counts[key] > best_count) {
      best = key;
      best_count = counts[key];
    }
  return best;
}

}  // namespace

extern "C" int svt_frame2note(const float* p_on, const float* p_off, const int32_t* oct, const int32_t* pc, int n_frames,
                              double onset_thres, double offset_thres, double frame_size, double* notes_out,
                              int max_notes, int* n_notes) {
  if (n_notes == nullptr || (n_frames > 0 && (p_on == nullptr || p_off == nullptr || oct == nullptr || pc == nullptr)))
    return svt::fail(svt::kInvalidArgument, "frame2note: null argument");
  const float on_thr = static_cast<float>(onset_thres);
  const float off_thr = static_cast<float>(offset_thres);
  int count = 0;
  bool open = false;
  double onset_time = 0.0, now = 0.0;
  std::vector<int> pitches;
  auto close_note = [&](double t_end) -> bool {
    if (pitches.empty()) return true;
    if (count >= max_notes || notes_out == nullptr) return false;
    double* row = notes_out + 3 * static_cast<size_t>(count++);
    row[0] = onset_time;
    row[1] = t_end;
    row[2] = static_cast<double>(pitch_mode(pitches) + 36);
    return true;
  };
  for (int i = 0; i < n_frames; ++i) {
    now = frame_size * static_cast<double>(i);
    bool onset = false;
    if (p_on[i] >= on_thr) {
      const int lo = i > 3 ? i - 3 : 0;
      const int hi = (i + 4 < n_frames - 1) ? i + 4 : n_frames - 1;
      if (hi <= lo) return svt::fail(svt::kInvalidArgument, "frame2note: empty local-max window (n_frames == 1), the reference raises here");
      float peak = p_on[lo];
      for (int j = lo + 1; j < hi; ++j) peak = p_on[j] > peak ? p_on[j] : peak;
      onset = (p_on[i] == peak);
    }
    if (onset) {
      if (open && !close_note(now)) return svt::fail(svt::kInvalidArgument, "frame2note: notes_out too small");
      open = true;
      onset_time = now;
      pitches.clear();
    } else if (p_off[i] >= off_thr && open) {
      if (!close_note(now)) return svt::fail(svt::kInvalidArgument, "frame2note: notes_out too small");
      open = false;
      pitches.clear();
    }
    if (open && oct[i] != 4 && pc[i] != 12) {
      const int pitch = oct[i] * 12 + pc[i];
      if (pitch < 0 || pitch >= 64) return svt::fail(svt::kInvalidArgument, "frame2note: pitch class out of range");
      pitches.push_back(pitch);
    }
  }
  if (open && !close_note(now)) return svt::fail(svt::kInvalidArgument, "frame2note: notes_out too small");
  *n_notes = count;
  return SVT_OK;
}
