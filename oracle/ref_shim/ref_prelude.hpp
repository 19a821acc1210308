// ref_prelude.hpp - force-included (-include) into every translation unit of the oracle/_ref build.
// The reference seeds its std::mt19937 generators from std::random_device (uniform_random_generator.hpp:11-14), which makes the
// PROSAC / LO draws unrepeatable. For the cross-checks the device is replaced by one that returns a seed chosen by the test
// driver; nothing else of the reference is altered. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <random>
inline unsigned& usac_ref_random_device_seed() { static unsigned seed = 5489u; return seed; }
namespace std {
struct usac_ref_fixed_random_device {
    unsigned operator()() { return usac_ref_random_device_seed(); }
};
}
#define random_device usac_ref_fixed_random_device
