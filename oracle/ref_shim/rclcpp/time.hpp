// TEST INFRASTRUCTURE ONLY (oracle/_ref build): the gait sequencer headers include rclcpp/time.hpp but use nothing of it.
#pragma once
namespace rclcpp { class Time {}; }
