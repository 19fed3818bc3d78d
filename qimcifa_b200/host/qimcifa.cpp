// qimcifa -- host driver with the reference's stdin dialog and output lines (reference src/qimcifa.cpp:61-207),
// sweeping on B200 GPUs through the C ABI.  One worker thread per GPU plays the role of the reference's
// hardware_concurrency() CPU workers (src/qimcifa.cpp:160-178): all pull batch ranges from the node's dispenser.
//
// Extra (optional) argv, for scripting; the dialog itself is unchanged:
//   --gpus G                 number of GPUs to use (default: all visible)
//   --batches-per-launch B   reference batches per kernel launch (default 1024)
//   --kernel auto|exact|filter
//   --literal-precheck       reproduce the snapshot's small-prime pre-check, which tests only every other
//                            wheel prime (src/qimcifa.cpp:134-142); default tests all of them
//   --level-10               accept wheel level 10 (prime 29), which the reference prompts for but clamps away
//   --gcd                    the reference's IS_RSA_SEMIPRIME=OFF build (CMakeLists.txt:7): report gcd(guess, N) != 1,
//                            "Has common factor: Found ..." (include/qimcifa.hpp:373-379)
//   --stats                  print guesses/s after the run
//   --random SEED            random-guess Monte-Carlo mode (reference README.md:25; no reference code exists):
//                            guesses are drawn by Philox4x32-10 through the wheel (qcf_sweep_random) instead of
//                            swept; GPU k of node `nodeId` takes disjoint counter blocks
//   --random-budget B        total guesses to draw in random mode (default 2^36), in rounds of 2^34 (chain draws) or 2^30 per GPU
//   --random-independent     one exact test per independently drawn guess instead of chain draws through the filter kernel
#include <cfloat>
#include <cstring>
#include <fstream>
#include <future>
#include <thread>

#include "qimcifa_host.hpp"

namespace Qimcifa {

struct Options {
    int gpus = 0;
    uint64_t batchesPerLaunch = 1024;
    uint32_t flags = QCF_KERNEL_AUTO;
    bool literalPrecheck = false;
    bool stats = false;
    bool gcd = false;
    bool randomMode = false;
    uint64_t randomSeed = 0;
    uint64_t randomBudget = 1ULL << 36;
    bool randomIndependent = false;
};

// "You can split this work across nodes" dialog: node count (>= 1), then this node's id when there are several.
static void askNodeSplit(size_t& nodeCount, size_t& nodeId)
{
    std::cout << "You can split this work across nodes, without networking!" << std::endl;
    for (nodeCount = 0U; !nodeCount;) {
        std::cout << "Number of nodes (>=1): ";
        std::cin >> nodeCount;
        if (!nodeCount) {
            std::cout << "Invalid node count choice!" << std::endl;
        }
    }
    nodeId = 0U;
    while (nodeCount > 1U) {
        std::cout << "Which node is this? (0-" << (nodeCount - 1U) << "): ";
        std::cin >> nodeId;
        if (nodeId < nodeCount) {
            break;
        }
        std::cout << "Invalid node ID choice!" << std::endl;
    }
}

// Level prompt and clamp: 0..4 -> 5, above 9 -> 9, negative -> calibration file (returned as -1).
static int64_t g_maxLevel = 9; // the reference clamps to 9 although it lists 10 primes (src/qimcifa.cpp:64,96-98); --level-10 lifts it

static int64_t askWheelLevel(size_t listedPrimes)
{
    int64_t level = 7; // what the reference's variable holds if the read fails: the "default wheel level"
    std::cout << "Wheel factorization level (minimum of " << MIN_RTD_LEVEL << ", max of " << listedPrimes
              << ", or -1 for calibration file): ";
    std::cin >> level;
    if (level < 0) {
        return -1;
    }
    return std::min<int64_t>(g_maxLevel, std::max<int64_t>(MIN_RTD_LEVEL, level));
}

// Picks the cheapest row of the tuner's file ("level cardinality batch_time cost" after one header line).
static int64_t readCalibration(const char* path, double& bestCost)
{
    std::ifstream settingsFile(path);
    if (!settingsFile.good()) {
        std::cerr << "qimcifa: " << path << " not found; run qimcifa_tuner first" << std::endl;
        return -1;
    }
    std::string header;
    std::getline(settingsFile, header);
    int64_t best = -1;
    size_t level;
    BigInteger cardinality;
    double batchTime, cost;
    while (settingsFile >> level >> cardinality >> batchTime >> cost) {
        if ((level >= (size_t)MIN_RTD_LEVEL) && (level <= 9U) && (cost < bestCost)) {
            best = (int64_t)level;
            bestCost = cost;
        }
    }
    if (best < 0) {
        std::cerr << "qimcifa: no usable row in " << path << std::endl;
    }
    return best;
}

int mainBody(const BigInteger& toFactor, const Options& opt)
{
    // First 9 primes (the reference lists 10 and clamps the level to 9, src/qimcifa.cpp:64,96-98)
    std::vector<unsigned> trialDivisionPrimes = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29 };

    // every return below releases the contexts (streams, events, tables, pinned records)
    struct ContextSet : std::vector<qcf_ctx*> {
        ~ContextSet()
        {
            activeContexts.clear();
            for (qcf_ctx* c : *this) {
                qcf_destroy(c);
            }
        }
    } contexts;
    int gpuCount = opt.gpus;
    for (int d = 0; (gpuCount == 0) || (d < gpuCount); ++d) {
        qcf_ctx* c = nullptr;
        if (qcf_create(d, &c) != QCF_OK) {
            if (gpuCount != 0 || d == 0) {
                std::cerr << "qimcifa: " << qcf_last_error() << std::endl;
                return 1;
            }
            break;
        }
        contexts.push_back(c);
    }
    const unsigned cpuCount = (unsigned)contexts.size(); // workers = GPUs

    // ---- the reference's dialog (src/qimcifa.cpp:70-123), same prompts in the same order ----
    size_t nodeCount = 1U, nodeId = 0U;
    askNodeSplit(nodeCount, nodeId);
    int64_t tdLevel = askWheelLevel(trialDivisionPrimes.size());
    if (tdLevel < 0) {
        double bestCost = DBL_MAX;
        tdLevel = readCalibration("qimcifa_calibration.ssv", bestCost);
        if (tdLevel < 0) {
            return 1; // the reference indexes out of bounds when the file is missing; we stop instead
        }
        std::cout << "Calibrated wheel factorization level: " << tdLevel << std::endl;
        std::cout << "Estimated average time to exit: " << (bestCost / (2 * cpuCount * nodeCount)) << " seconds" << std::endl;
    }

    // Starting clock right after user input is finished
    auto iterClock = std::chrono::high_resolution_clock::now();

    // pre-checks, with the reference's output lines (src/qimcifa.cpp:128-142): perfect square, then the wheel's own primes
    // (the snapshot advances its index twice per round and so tests every other prime, SURVEY.md D4: --literal-precheck)
    uint32_t n_le[QCF_MAX_N_LIMBS], sqrt_le[QCF_MAX_N_LIMBS];
    toFactor.toLimbs(n_le, QCF_MAX_N_LIMBS);
    uint64_t nodeRange64 = 0, batchCount64 = 0;
    if (qcf_node_split(n_le, QCF_MAX_N_LIMBS, nodeCount, sqrt_le, nullptr, &nodeRange64, &batchCount64)) {
        std::cerr << "qimcifa: " << qcf_last_error() << std::endl; // includes "batch count exceeds 64 bits"
        return 1;
    }
    const BigInteger root = BigInteger::fromLimbs(sqrt_le, QCF_MAX_N_LIMBS);
    if (root * root == toFactor) {
        std::cout << "Number to factor is a perfect square: " << root << " * " << root << " = " << toFactor;
        return 0;
    }
    for (int64_t i = 0; i < tdLevel; i += opt.literalPrecheck ? 2 : 1) {
        const unsigned prime = trialDivisionPrimes[(size_t)i];
        if (!(bool)(toFactor % prime)) {
            std::cout << "Factors: " << prime << " * " << (toFactor / prime) << " = " << toFactor << std::endl;
            return 0;
        }
    }

    // the node's slice of the dispenser: qcf_node_split is the library's statement of src/qimcifa.cpp:149,155-158
    // (fullRange = backward(floor sqrt N), nodeRange batches per node); node k hands out batches descending from the top of
    // the k-th slice from the top
    const BigInteger nodeRange(nodeRange64);
    batchCount = BigInteger(batchCount64);
    batchNumber = nodeRange * BigInteger((uint64_t)nodeId);
    batchBound = batchNumber + nodeRange;
    activeContexts = contexts;

    std::vector<SweepTotals> totals(cpuCount);
    std::vector<std::future<void>> futures;
    futures.reserve(cpuCount);
    for (unsigned cpu = 0U; cpu < cpuCount; ++cpu) {
        if (opt.randomMode) {
            // Monte-Carlo mode: round r of worker w (w = nodeId*GPUs + gpu, W = nodeCount*GPUs workers) draws counters
            // [(r*W + w) * 2^30, +2^30): disjoint across GPUs and nodes, reproducible from (seed, nodeCount, nodeId).
            futures.push_back(std::async(std::launch::async, [&, cpu] {
                uint32_t n_le[QCF_MAX_N_LIMBS];
                toFactor.toLimbs(n_le, QCF_MAX_N_LIMBS);
                // chain draws (QCF_RANDOM_CHAINS: Philox picks row groups of chains of a random window, the filter kernel walks
                // them; counter blocks of 2^34) unless --random-independent or the number is too small for the filter kernel
                bool chainDraws = !opt.randomIndependent;
                uint64_t block = chainDraws ? (1ULL << 34) : (1ULL << 30);
                const uint64_t workers = (uint64_t)nodeCount * cpuCount, me = (uint64_t)nodeId * cpuCount + cpu;
                uint64_t rounds = (opt.randomBudget + block * workers - 1U) / (block * workers);
                for (uint64_t r = 0; r < rounds; ++r) {
                    {
                        std::lock_guard<std::mutex> lock(batchMutex);
                        if (batchNumber >= batchBound) {
                            return; // finish() was called
                        }
                    }
                    qcf_result res;
                    int rcRandom = qcf_sweep_random(contexts[cpu], n_le, QCF_MAX_N_LIMBS, (int)tdLevel, opt.randomSeed, (r * workers + me) * block, block,
                        chainDraws ? QCF_RANDOM_CHAINS : 0u, &res);
                    if ((rcRandom == QCF_ERANGE) && chainDraws) {
                        chainDraws = false; // too small for the filter kernel: independent draws from here on, same budget
                        block = 1ULL << 30;
                        rounds = (opt.randomBudget + block * workers - 1U) / (block * workers);
                        r = (uint64_t)0 - 1U;
                        continue;
                    }
                    if (rcRandom) {
                        std::cerr << "qcf_sweep_random failed: " << qcf_last_error() << std::endl;
                        finish();
                        return;
                    }
                    totals[cpu].guesses += res.guesses_tested;
                    totals[cpu].launches += res.kernel_launches;
                    totals[cpu].deviceSeconds += res.device_seconds;
                    if (res.found) {
                        const BigInteger base = BigInteger::fromLimbs(res.factor_le, QCF_MAX_N_LIMBS);
                        printSuccess(base, toFactor / base, toFactor, "Exact factor: Found ", iterClock);
                        return;
                    }
                }
            }));
            continue;
        }
        futures.push_back(std::async(std::launch::async, [&, cpu] {
            qcf_build_tables(contexts[cpu], (int)tdLevel); // device-side wheel generation (replaces wheel_gen, :152)
            getSmoothNumbers(contexts[cpu], toFactor, (int)tdLevel, opt.batchesPerLaunch, opt.flags, iterClock, &totals[cpu]);
        }));
    }
    for (unsigned cpu = 0U; cpu < cpuCount; ++cpu) {
        futures[cpu].get();
    }

    if (opt.stats) {
        SweepTotals t;
        for (const SweepTotals& s : totals) {
            t.guesses += s.guesses;
            t.batches += s.batches;
            t.launches += s.launches;
            t.deviceSeconds = std::max(t.deviceSeconds, s.deviceSeconds);
        }
        const double wall = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - iterClock).count() * 1e-6;
        std::cout << "(Stats: " << t.guesses << " guesses in " << t.batches << " batches, " << t.launches << " launches, " << cpuCount
                  << " GPU(s), " << wall << " s wall, " << (t.guesses / wall) << " guesses/s)" << std::endl;
    }
    return 0;
}
} // namespace Qimcifa

using namespace Qimcifa;

int main(int argc, char** argv)
{
    Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if ((a == "--gpus") && (i + 1 < argc)) {
            opt.gpus = atoi(argv[++i]);
        } else if ((a == "--batches-per-launch") && (i + 1 < argc)) {
            opt.batchesPerLaunch = strtoull(argv[++i], nullptr, 10);
        } else if ((a == "--kernel") && (i + 1 < argc)) {
            const std::string k = argv[++i];
            opt.flags = (k == "exact") ? QCF_KERNEL_EXACT : ((k == "filter") ? QCF_KERNEL_FILTER : QCF_KERNEL_AUTO);
        } else if (a == "--literal-precheck") {
            opt.literalPrecheck = true;
        } else if (a == "--level-10") {
            g_maxLevel = 10;
        } else if (a == "--gcd") {
            opt.gcd = true;
        } else if (a == "--stats") {
            opt.stats = true;
        } else if ((a == "--random") && (i + 1 < argc)) {
            opt.randomMode = true;
            opt.randomSeed = strtoull(argv[++i], nullptr, 10);
        } else if ((a == "--random-budget") && (i + 1 < argc)) {
            opt.randomBudget = strtoull(argv[++i], nullptr, 10);
        } else if (a == "--random-independent") {
            opt.randomIndependent = true;
        } else {
            std::cerr << "usage: qimcifa [--gpus G] [--batches-per-launch B] [--kernel auto|exact|filter] [--literal-precheck] [--gcd] [--level-10] [--stats] [--random SEED] [--random-budget B] [--random-independent]" << std::endl;
            return 2;
        }
    }
    if (opt.gcd) {
        opt.flags |= QCF_MODE_GCD;
    }
    if (!opt.batchesPerLaunch) {
        opt.batchesPerLaunch = 1024;
    }

    BigInteger toFactor;
    std::cout << "Number to factor: ";
    std::cin >> toFactor;

    // the banner lines of the reference (src/qimcifa.cpp:193-204): a power of two 2^k is announced and counted as k bits,
    // anything else as its bit length
    const bool powerOfTwo = isPowerOfTwo(toFactor);
    if (powerOfTwo) {
        std::cout << "Number to factor is a power of 2: 2 * " << (toFactor >> 1U) << " = " << toFactor;
    }
    std::cout << "Bits to factor: " << (toFactor.bitLength() - (powerOfTwo ? 1 : 0)) << std::endl;
    if (toFactor.bitLength() > 32 * QCF_MAX_N_LIMBS) {
        std::cerr << "qimcifa: numbers above " << 32 * QCF_MAX_N_LIMBS << " bits are not supported by the GPU path" << std::endl;
        return 1;
    }
    if (toFactor < 4U) {
        std::cerr << "qimcifa: nothing to factor" << std::endl;
        return 1;
    }

    return mainBody(toFactor, opt);
}
