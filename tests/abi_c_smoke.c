/* Plain C (not C++) user of include/qimcifa_cuda.h: proves the header is a C ABI.  Host-only calls. */
#include <stdio.h>
#include <string.h>

#include "qimcifa_cuda.h"

int main(void)
{
    /* N = 18446524746962277769 = 4294917283 * 4294966243 */
    const uint32_t n_le[4] = { 0x03239589u, 0xFFFF3886u, 0u, 0u };
    uint32_t root[4], range[4], count[4];
    uint64_t node_range = 0, batch_count = 0;
    qcf_ctx* ctx = 0;
    int rc = qcf_node_split(n_le, 4, 8, root, range, &node_range, &batch_count);
    if (rc != QCF_OK) {
        printf("node_split failed: %s\n", qcf_last_error());
        return 1;
    }
    printf("sqrt=%u nodeRange=%llu batchCount=%llu\n", root[0], (unsigned long long)node_range, (unsigned long long)batch_count);
    rc = qcf_count_guesses(5, 0, 1, count);
    printf("count=%u rc=%d\n", count[0], rc);
    rc = qcf_node_split(n_le, 0, 1, root, range, &node_range, &batch_count); /* bad argument -> error code, no abort */
    printf("bad=%d msg=%s\n", rc, qcf_last_error());
    rc = qcf_create(0, &ctx);
    printf("create=%d\n", rc);
    if (rc == QCF_OK) {
        qcf_destroy(ctx);
    }
    return 0;
}
