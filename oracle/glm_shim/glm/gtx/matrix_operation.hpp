// TEST INFRASTRUCTURE ONLY: see ../glm.hpp (minimal glm-compatible shim).
#pragma once
#include "../glm.hpp"
