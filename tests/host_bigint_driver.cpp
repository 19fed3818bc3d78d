// Test driver for qimcifa_b200/host: reads "op a b" lines, prints the result computed with the host
// BigInteger / qimcifa_host.hpp helpers (compared against Python ints by tests/test_host.py).
#include <iostream>
#include <string>

#include "qimcifa_host.hpp"

using namespace Qimcifa;

int main()
{
    std::string op;
    BigInteger a, b;
    while (std::cin >> op >> a >> b) {
        if (op == "add") {
            std::cout << (a + b);
        } else if (op == "sub") {
            std::cout << (a - b);
        } else if (op == "mul") {
            std::cout << (a * b);
        } else if (op == "div") {
            std::cout << (a / b);
        } else if (op == "mod") {
            std::cout << (a % b);
        } else if (op == "and") {
            std::cout << (a & b);
        } else if (op == "shr") {
            std::cout << (a >> (unsigned)(size_t)b);
        } else if (op == "shl") {
            std::cout << (a << (unsigned)(size_t)b);
        } else if (op == "cmp") {
            std::cout << (a < b) << (a <= b) << (a > b) << (a >= b) << (a == b) << (a != b);
        } else if (op == "sqrt") {
            std::cout << sqrt(a);
        } else if (op == "fwd") {
            std::cout << forward(a);
        } else if (op == "bwd") {
            std::cout << backward(a);
        } else if (op == "pow2") {
            std::cout << isPowerOfTwo(a);
        } else if (op == "mixed") { // built-in operands on both sides, as the drivers use them
            std::cout << ((size_t)3 * a + 7U - 1U) / (size_t)5 % 1000003U;
        } else if (op == "dbl") {
            std::cout << (uint64_t)(a.convert_to<double>() / 1024.0);
        } else if (op == "split") { // src/qimcifa.cpp:149,155-158 ; b = nodeCount
            const size_t nodeCount = (size_t)b;
            const BigInteger fullRange = backward(sqrt(a));
            const BigInteger nodeRange = (((fullRange + nodeCount - 1U) / nodeCount) + BIGGEST_WHEEL - 1U) / BIGGEST_WHEEL;
            std::cout << fullRange << " " << nodeRange << " " << (nodeCount * nodeRange);
        } else if (op == "disp") { // dispenser: a = nodeRange, b = nodeCount*16 + nodeId
            const size_t nodeCount = (size_t)b / 16, nodeId = (size_t)b % 16;
            batchNumber = nodeId * a;
            batchBound = (nodeId + 1) * a;
            batchCount = nodeCount * a;
            uint64_t lo, hi;
            while (getNextBatches(2, lo, hi)) {
                std::cout << lo << ":" << hi << ",";
            }
        }
        std::cout << "\n";
    }
    return 0;
}
