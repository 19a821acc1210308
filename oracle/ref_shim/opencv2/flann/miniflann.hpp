// stand-in for <opencv2/flann/miniflann.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../../cvshim.hpp"
