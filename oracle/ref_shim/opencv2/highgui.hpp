// stand-in for <opencv2/highgui.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
#include "../cvshim.hpp"
