#pragma once
// forwarder: <pico_understory/kd_forest.hpp> of the reference -> the device-backed class
#include "../../pico_tree_b200/kd_forest.hpp"
