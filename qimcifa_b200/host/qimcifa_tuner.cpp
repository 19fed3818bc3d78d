// qimcifa_tuner -- same dialog and qimcifa_calibration.ssv format as the reference tuner
// (reference src/qimcifa_tuner.cpp:100-181), but it MEASURES: per wheel level it times real batches just below
// sqrt(N) on the GPU (qcf_time_level) and writes
//     level cardinality batch_time(s) cost(s)
// with cost = turboCorrection * batch_time * batchCount, exactly the reference's cost model (:139-143); the
// reference's own timing loop times nothing in this snapshot.  turboCorrection defaults to 1 here
// (the GPU runs at a steady clock; the reference's default 2e7 compensates for its empty timing loop).
//   --gpu D            device to time on (default 0)
//   --batches B        batches timed per level (default 1024: the launch shape of the sweep)
//   --kernel auto|exact|filter
#include <cfloat>
#include <cstring>
#include <fstream>

#include "qimcifa_host.hpp"

using namespace Qimcifa;

int main(int argc, char** argv)
{
    int device = 0;
    uint64_t batches = 1024;
    uint32_t flags = QCF_KERNEL_AUTO;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if ((a == "--gpu") && (i + 1 < argc)) {
            device = atoi(argv[++i]);
        } else if ((a == "--batches") && (i + 1 < argc)) {
            batches = strtoull(argv[++i], nullptr, 10);
        } else if ((a == "--kernel") && (i + 1 < argc)) {
            const std::string k = argv[++i];
            flags = (k == "exact") ? QCF_KERNEL_EXACT : ((k == "filter") ? QCF_KERNEL_FILTER : QCF_KERNEL_AUTO);
        } else {
            std::cerr << "usage: qimcifa_tuner [--gpu D] [--batches B] [--kernel auto|exact|filter]" << std::endl;
            return 2;
        }
    }

    BigInteger toFactor;
    std::cout << "The qimcifa_tuner number need not be the exact number that will be factored with qimcifa, but closer is better." << std::endl;
    std::cout << "Number to factor: ";
    std::cin >> toFactor;

    uint32_t qubitCount = 0;
    BigInteger p = toFactor >> 1U;
    while (p) {
        p >>= 1U;
        ++qubitCount;
    }
    if (!isPowerOfTwo(toFactor)) {
        qubitCount++;
    }
    std::cout << "Bits to factor: " << (int)qubitCount << std::endl;

    size_t threadCount = 1;
    std::cout << "Total thread count (across all nodes): ";
    std::cin >> threadCount;
    if (!threadCount) {
        threadCount = 1;
    }
    double turboCorrection = 1.0;
    std::cout << "Turbo correction multiplier (0 for default): ";
    std::cin >> turboCorrection;
    if (turboCorrection <= 0) {
        turboCorrection = 1.0;
    }
    if ((toFactor.bitLength() > 32 * QCF_MAX_N_LIMBS) || (toFactor < 4U)) {
        std::cerr << "qimcifa_tuner: unsupported number" << std::endl;
        return 1;
    }

    qcf_ctx* ctx = nullptr;
    if (qcf_create(device, &ctx) != QCF_OK) {
        std::cerr << "qimcifa_tuner: " << qcf_last_error() << std::endl;
        return 1;
    }
    uint32_t n_le[QCF_MAX_N_LIMBS];
    toFactor.toLimbs(n_le, QCF_MAX_N_LIMBS);

    std::vector<unsigned> trialDivisionPrimes = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29 };

    const BigInteger root = sqrt(toFactor);
    const BigInteger batchCount = (backward(root) + BIGGEST_WHEEL - 1) / BIGGEST_WHEEL; // batches of the whole range
    BigInteger range = backward(root);                                                  // reference :133
    std::ofstream oSettingsFile("qimcifa_calibration.ssv");
    oSettingsFile << "level, cardinality, batch time (s), cost (s)" << std::endl;
    for (size_t i = MIN_RTD_LEVEL; i < 10U; ++i) {
        double seconds = 0;
        uint64_t guesses = 0;
        const uint64_t timed = (batchCount < BigInteger(batches)) ? (uint64_t)batchCount : batches;
        qcf_time_level(ctx, n_le, QCF_MAX_N_LIMBS, (int)i, timed, flags, &seconds, &guesses); // warm-up (tables, clocks)
        if (qcf_time_level(ctx, n_le, QCF_MAX_N_LIMBS, (int)i, timed, flags, &seconds, &guesses) != QCF_OK) {
            std::cerr << "qimcifa_tuner: " << qcf_last_error() << std::endl;
            return 1;
        }
        const double time = seconds / (double)timed; // one batch
        oSettingsFile << i << " " << range << " " << time << " " << (turboCorrection * time * batchCount.convert_to<double>()) << std::endl;
        if (i < 9U) {
            const unsigned nextPrime = trialDivisionPrimes[i];
            range *= nextPrime - 1U;
            range = (range + nextPrime - 1U) / nextPrime; // reference :145-148
        }
    }
    oSettingsFile.close();
    qcf_destroy(ctx);

    std::ifstream iSettingsFile("qimcifa_calibration.ssv");
    std::string header;
    std::getline(iSettingsFile, header);
    size_t bestLevel = (size_t)-1;
    double bestCost = DBL_MAX;
    size_t level;
    BigInteger cardinality;
    double batchTime, cost;
    while (iSettingsFile >> level >> cardinality >> batchTime >> cost) {
        if ((level >= (size_t)MIN_RTD_LEVEL) && (cost < bestCost)) {
            bestLevel = level;
            bestCost = cost;
        }
    }
    iSettingsFile.close();

    std::cout << "Calibrated reverse trial division level: " << bestLevel << std::endl;
    std::cout << "Estimated average time to exit: " << (bestCost / (2 * threadCount)) << " seconds" << std::endl;

    return 0;
}
