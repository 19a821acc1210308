// stand-in for <opencv2/flann/flann.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../../cvshim.hpp"
