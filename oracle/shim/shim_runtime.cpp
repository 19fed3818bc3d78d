// Link-time override of std::thread::hardware_concurrency (TEST INFRASTRUCTURE ONLY).
// The reference spawns hardware_concurrency() workers (reference src/qimcifa.cpp:66,172-178).
// To obtain the single-thread literal trace (SURVEY.md Appendix A, S-literal) from the
// UNMODIFIED sources, the environment variable QCF_REF_THREADS selects the worker count;
// unset, the real core count is used as in the reference.
#include <cstdlib>
#include <thread>
#include <unistd.h>

unsigned int std::thread::hardware_concurrency() noexcept
{
    const char* e = getenv("QCF_REF_THREADS");
    if (e && *e) {
        const int v = atoi(e);
        if (v > 0) {
            return (unsigned)v;
        }
    }
    const long n = sysconf(_SC_NPROCESSORS_ONLN);
    return (n > 0) ? (unsigned)n : 1U;
}
