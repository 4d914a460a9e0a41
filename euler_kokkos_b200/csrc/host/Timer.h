// Accumulating wall-clock timer with the reference's start/stop/elapsed interface
// (src/utils/monitoring/OpenMPTimer.h, CudaTimer.h).  Device work is asynchronous here, so timers
// that bracket un-synchronised calls measure enqueue time only; TIMER_TOTAL brackets a synchronised
// region (see SolverBase).
#pragma once
#include <chrono>

namespace euler_b200 {

class Timer {
public:
  void start() { t0_ = clock::now(); }
  void stop() { total_ += std::chrono::duration<double>(clock::now() - t0_).count(); }
  double elapsed() const { return total_; }
  void reset() { total_ = 0.0; }

private:
  using clock = std::chrono::steady_clock;
  clock::time_point t0_{};
  double total_ = 0.0;
};

} // namespace euler_b200
