// ref_math_wrap.cpp — TEST INFRASTRUCTURE (oracle), never linked into the product.
//
// Compiles the reference's own Src/Math.cpp, unmodified and in place, with exactly one behaviour
// replaced: RMath::PseudoRandomUnitVector (Math.cpp:33-40) walks its 16.7 M-entry table through
// a single function-local static cursor shared by every thread, so which direction a diffuse
// bounce gets depends on thread scheduling.  Here the table index is drawn from rand() instead
// (which the harness interposes with the counter RNG of include/rt_rng.h), so a bounce's
// direction is a pure function of (seed, pixel, sample, draw#).  The table contents, the
// hemisphere flip (Math.cpp:42-54) and Barycentric (Math.cpp:56-82) are the reference's code.
//
// Mechanism: the token `PseudoRandomUnitVector` occurs twice in Math.cpp — its definition and
// its call inside RandomHemisphereDirection.  A __COUNTER__-suffixed rename sends the first to
// an unused symbol and the second to the replacement defined below (which can see the
// anonymous-namespace table because it lives in the same translation unit).
#include <stdlib.h>
#include <fstream>
#include <mutex>
#include <string>
#include <math.h>

#include "Math.h"      // declares RMath::PseudoRandomUnitVector under its real name first
#include "Platform.h"

namespace RMath
{
    RVec3 RtOraclePRUV_1();   // the shipped cursor-based body (unused)
    RVec3 RtOraclePRUV_2();   // counter-RNG replacement, defined below
}

#define RT_ORACLE_CAT2(a, b) a##b
#define RT_ORACLE_CAT(a, b) RT_ORACLE_CAT2(a, b)
static_assert(__COUNTER__ == 0, "__COUNTER__ must start at 0 here");   // consumes value 0
#define PseudoRandomUnitVector RT_ORACLE_CAT(RtOraclePRUV_, __COUNTER__)
#include "Math.cpp"
#undef PseudoRandomUnitVector
static_assert(__COUNTER__ == 3, "expected exactly two occurrences of PseudoRandomUnitVector in Math.cpp");

namespace RMath
{
    RVec3 RtOraclePRUV_2()
    {
        return PseudoRandomUnitVectors[(unsigned int)rand() % MaxUnitVectorNums];
    }

    // keep the declared symbol defined for any other caller
    RVec3 PseudoRandomUnitVector()
    {
        return RtOraclePRUV_2();
    }
}

extern "C" const float* ref_unit_vector_table(unsigned int* count)
{
    *count = MaxUnitVectorNums;
    return &PseudoRandomUnitVectors[0].x;
}
