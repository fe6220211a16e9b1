/*
 * tests/election/harness.c -- runs the PRODUCT's election policy (apus_b200/csrc/dare_entry.c: elect, included here as
 * source) in one survivor process of a group whose leader has just died, on the mock control words of mock_engine.c.
 *   harness <dir> <idx> <n> <dead> <term> <last_idx> <last_term> <commit> <end> <elec_lo_us> <elec_hi_us>
 * prints one line:  RESULT idx=<i> role=<leader|follower|none> leader=<l> term=<t>
 */
#include "../../apus_b200/csrc/dare_entry.c"

apus_replica_t *mock_open(const char *dir, int idx, int n, uint64_t last_idx, uint64_t last_term, uint64_t commit, uint64_t end, uint64_t sid);

int main(int argc, char **argv)
{
    if (argc < 12) return 2;
    const char *dir = argv[1];
    g_idx = (uint8_t)atoi(argv[2]); g_n = (uint8_t)atoi(argv[3]); g_leader_idx = (uint8_t)atoi(argv[4]);
    g_term = strtoull(argv[5], NULL, 0);
    g_live_mask = (1u << g_n) - 1u;
    g_log = stdout;
    g_log_len = APUS_LOG_SIZE;
    cfg_hb_period = 0.002;
    cfg_elec_low = strtoull(argv[10], NULL, 0); cfg_elec_high = strtoull(argv[11], NULL, 0);
    snprintf(g_env_rdv, sizeof g_env_rdv, "%s/rdv", dir);
    g_rep = mock_open(dir, g_idx, g_n, strtoull(argv[6], NULL, 0), strtoull(argv[7], NULL, 0), strtoull(argv[8], NULL, 0),
                      strtoull(argv[9], NULL, 0), SID_MAKE(g_term, 1, g_leader_idx));
    /* every survivor is up before anybody stands (the drill starts them together) */
    char path[600];
    snprintf(path, sizeof path, "%s/up%u", dir, (unsigned)g_idx);
    FILE *f = fopen(path, "w"); if (f) fclose(f);
    for (int tries = 0; tries < 5000; tries++) {
        int all = 1;
        for (unsigned i = 0; i < g_n; i++) { if (i == g_leader_idx) continue; snprintf(path, sizeof path, "%s/up%u", dir, i); if (access(path, F_OK)) all = 0; }
        if (all) break;
        usleep(1000);
    }
    const uint8_t dead = g_leader_idx;
    int rc = elect();
    printf("RESULT idx=%u role=%s leader=%u term=%llu rc=%d dead=%u\n", (unsigned)g_idx,
           rc ? "none" : (g_leader_idx == g_idx ? "leader" : "follower"), (unsigned)g_leader_idx, (unsigned long long)g_term, rc, (unsigned)dead);
    fflush(stdout);
    /* a new leader keeps its words alive until the followers have read them */
    usleep(300000);
    return rc;
}
