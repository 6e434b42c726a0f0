// TEST INFRASTRUCTURE ONLY (oracle/_ref build): the fields of the ROS 2 message interfaces/msg/GaitState that
// SimpleGaitSequencer::GetGaitState / AdaptiveGaitSequencer::GetGaitState write
// (simple_gait_sequencer.cpp:84-101, adaptive_gait_sequencer.cpp:64-71).
#pragma once
#include <array>
#include <cstdint>
namespace interfaces { namespace msg {
struct GaitState {
  double period = 0.0;
  std::array<double, 4> duty_factor{}, phase_offset{}, phase{};
  std::array<bool, 4> contact{};
  uint8_t gait_sequencer = 0;
};
} }
