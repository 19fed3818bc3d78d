// qimcifa_tuner -- same dialog and qimcifa_calibration.ssv format as the reference tuner
// (reference src/qimcifa_tuner.cpp:100-181), but it MEASURES: per wheel level it times real batches just below
// sqrt(N) on the GPU (qcf_time_level) and writes
//     level cardinality batch_time(s) cost(s)
// with cost = turboCorrection * batch_time * batchCount, the reference's cost model (:139-143); the reference's own
// timing loop times nothing in this snapshot.  turboCorrection defaults to 1 here (the GPU runs at a
// steady clock; the reference's default 2e7 compensates for its empty timing loop).
//   --gpu D            device to time on (default 0)
//   --batches B        batches timed per level (default 1024: the launch shape of the sweep)
//   --kernel auto|exact|filter
#include <cfloat>
#include <cstring>
#include <fstream>

#include "qimcifa_host.hpp"

using namespace Qimcifa;

namespace {

const char* kCalibrationFile = "qimcifa_calibration.ssv";

struct TunerOptions {
    int device = 0;
    uint64_t batches = 1024;
    uint32_t flags = QCF_KERNEL_AUTO;
};

bool parseArgs(int argc, char** argv, TunerOptions& opt)
{
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const bool hasValue = i + 1 < argc;
        if ((a == "--gpu") && hasValue) {
            opt.device = atoi(argv[++i]);
        } else if ((a == "--batches") && hasValue) {
            opt.batches = strtoull(argv[++i], nullptr, 10);
        } else if ((a == "--kernel") && hasValue) {
            const std::string k = argv[++i];
            opt.flags = (k == "exact") ? QCF_KERNEL_EXACT : ((k == "filter") ? QCF_KERNEL_FILTER : QCF_KERNEL_AUTO);
        } else {
            return false;
        }
    }
    return opt.batches > 0;
}

// "Bits to factor" as the reference counts them (src/qimcifa_tuner.cpp:107-116)
int bitsToFactor(const BigInteger& n)
{
    int bits = 0;
    for (BigInteger p = n >> 1U; p; p >>= 1U) {
        ++bits;
    }
    return bits + (isPowerOfTwo(n) ? 0 : 1);
}

// One row per level 5..9; the cardinality column follows the reference's recurrence range = ceil(range*(p-1)/p).
int writeCalibration(qcf_ctx* ctx, const BigInteger& toFactor, const TunerOptions& opt, double turboCorrection)
{
    static const unsigned nextPrimeOfLevel[10] = { 0, 0, 0, 0, 0, 13, 17, 19, 23, 29 };
    uint32_t n_le[QCF_MAX_N_LIMBS];
    toFactor.toLimbs(n_le, QCF_MAX_N_LIMBS);
    const BigInteger root = sqrt(toFactor);
    BigInteger cardinality = backward(root);
    const BigInteger batchCount = (cardinality + BIGGEST_WHEEL - 1) / BIGGEST_WHEEL; // batches of the whole range
    const uint64_t timed = (batchCount < BigInteger(opt.batches)) ? (uint64_t)batchCount : opt.batches;
    std::ofstream out(kCalibrationFile);
    out << "level, cardinality, batch time (s), cost (s)" << std::endl;
    for (int level = MIN_RTD_LEVEL; level <= 9; ++level) {
        double seconds = 0;
        uint64_t guesses = 0;
        qcf_time_level(ctx, n_le, QCF_MAX_N_LIMBS, level, timed, opt.flags, &seconds, &guesses); // warm-up: tables, clocks
        if (qcf_time_level(ctx, n_le, QCF_MAX_N_LIMBS, level, timed, opt.flags, &seconds, &guesses) != QCF_OK) {
            std::cerr << "qimcifa_tuner: " << qcf_last_error() << std::endl;
            return 1;
        }
        const double batchTime = seconds / (double)timed;
        out << level << " " << cardinality << " " << batchTime << " " << (turboCorrection * batchTime * batchCount.convert_to<double>())
            << std::endl;
        const unsigned q = nextPrimeOfLevel[level];
        cardinality = (cardinality * (q - 1U) + q - 1U) / q;
    }
    return 0;
}

// Re-reads the file the way `qimcifa` does with level -1 and returns the cheapest level.
size_t bestLevelOfFile(double& bestCost)
{
    std::ifstream in(kCalibrationFile);
    std::string header;
    std::getline(in, header);
    size_t bestLevel = (size_t)-1, level;
    BigInteger cardinality;
    double batchTime, cost;
    bestCost = DBL_MAX;
    while (in >> level >> cardinality >> batchTime >> cost) {
        if ((level >= (size_t)MIN_RTD_LEVEL) && (cost < bestCost)) {
            bestLevel = level;
            bestCost = cost;
        }
    }
    return bestLevel;
}

} // namespace

int main(int argc, char** argv)
{
    TunerOptions opt;
    if (!parseArgs(argc, argv, opt)) {
        std::cerr << "usage: qimcifa_tuner [--gpu D] [--batches B] [--kernel auto|exact|filter]" << std::endl;
        return 2;
    }

    BigInteger toFactor;
    std::cout << "The qimcifa_tuner number need not be the exact number that will be factored with qimcifa, but closer is better." << std::endl;
    std::cout << "Number to factor: ";
    std::cin >> toFactor;
    std::cout << "Bits to factor: " << bitsToFactor(toFactor) << std::endl;

    size_t threadCount = 1;
    std::cout << "Total thread count (across all nodes): ";
    std::cin >> threadCount;
    double turboCorrection = 1.0;
    std::cout << "Turbo correction multiplier (0 for default): ";
    std::cin >> turboCorrection;
    if (turboCorrection <= 0) {
        turboCorrection = 1.0;
    }
    if (!threadCount) {
        threadCount = 1;
    }
    if ((toFactor.bitLength() > 32 * QCF_MAX_N_LIMBS) || (toFactor < 4U)) {
        std::cerr << "qimcifa_tuner: unsupported number" << std::endl;
        return 1;
    }

    qcf_ctx* ctx = nullptr;
    if (qcf_create(opt.device, &ctx) != QCF_OK) {
        std::cerr << "qimcifa_tuner: " << qcf_last_error() << std::endl;
        return 1;
    }
    const int rc = writeCalibration(ctx, toFactor, opt, turboCorrection);
    qcf_destroy(ctx);
    if (rc) {
        return rc;
    }

    double bestCost = DBL_MAX;
    const size_t bestLevel = bestLevelOfFile(bestCost);
    std::cout << "Calibrated reverse trial division level: " << bestLevel << std::endl;
    std::cout << "Estimated average time to exit: " << (bestCost / (2 * threadCount)) << " seconds" << std::endl;
    return 0;
}
