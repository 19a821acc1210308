// stand-in for <opencv2/core/mat.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../../cvshim.hpp"
