// Forwarder: lets sources written for the reference's <pico_tree/map_traits.hpp> compile against
// pico_tree_b200 when the reference's headers are not installed (add -I include/compat).
#pragma once
#include "../../pico_tree_b200/kd_tree.hpp"
