// TEST INFRASTRUCTURE ONLY (oracle/_ref build): potato_model.hpp includes rclcpp/rclcpp.hpp but uses nothing of it.
#pragma once
#include "rclcpp/time.hpp"
