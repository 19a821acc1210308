// stand-in for <opencv2/core/eigen.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../../cvshim.hpp"
