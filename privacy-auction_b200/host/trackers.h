// Wall-clock and byte accounting per category with the interface of the
// reference's include/timeTracker.h and include/dataTracker.h (observability
// only).  Unlike the reference, intervals are kept in microseconds and not
// truncated to whole milliseconds.
#ifndef PA_HOST_TRACKERS_H
#define PA_HOST_TRACKERS_H

#include <chrono>
#include <cstddef>
#include <iostream>
#include <map>
#include <string>

class TimeTracker {
public:
  static TimeTracker &getInstance() {
    static TimeTracker t;
    return t;
  }
  void start(const std::string &category) {
    auto &s = slots_[category];
    if (s.running) {
      std::cerr << "Timer for category " << category << " is already running.\n";
      return;
    }
    s.t0 = std::chrono::steady_clock::now();
    s.running = true;
  }
  void stop(const std::string &category) {
    auto &s = slots_[category];
    if (!s.running) {
      std::cerr << "Timer for category " << category << " is not running.\n";
      return;
    }
    s.total_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - s.t0).count();
    s.running = false;
  }
  double getCategoryTimeInSeconds(const std::string &category) const {
    auto it = slots_.find(category);
    return it == slots_.end() ? 0.0 : it->second.total_us / 1e6;
  }

private:
  struct Slot {
    std::chrono::steady_clock::time_point t0;
    long long total_us = 0;
    bool running = false;
  };
  std::map<std::string, Slot> slots_;
};

class DataTracker {
public:
  static DataTracker &getInstance() {
    static DataTracker t;
    return t;
  }
  void addData(const std::string &category, size_t size) {
    total_ += size;
    per_[category] += size;
  }
  size_t getTotalDataSize() const { return total_; }
  double getTotalDataSizeInMB() const { return (double)total_ / (1024 * 1024); }
  size_t getCategoryDataSize(const std::string &category) const {
    auto it = per_.find(category);
    return it == per_.end() ? 0 : it->second;
  }
  double getCategoryDataSizeInMB(const std::string &category) const { return (double)getCategoryDataSize(category) / (1024 * 1024); }

private:
  size_t total_ = 0;
  std::map<std::string, size_t> per_;
};

#endif
