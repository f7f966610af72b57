// oracle/shim: see opencv2/core.hpp (TEST INFRASTRUCTURE ONLY)
#pragma once
#include "core.hpp"
