// stand-in for <opencv2/features2d.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../cvshim.hpp"
