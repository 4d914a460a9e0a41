#include "comm_rendezvous.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <unistd.h>

#include "../../../include/ek_b200.h"

namespace euler_b200 {

static std::string g_rdv_path;

// called once ek_comm_attach has returned: ncclCommInitRank is collective, so every rank has read the id by then
void rendezvous_done(int rank)
{
  if (rank == 0 && !g_rdv_path.empty())
    std::remove(g_rdv_path.c_str());
  g_rdv_path.clear();
}

void rendezvous_unique_id(int rank, int nranks, unsigned char* id128)
{
  std::string path;
  if (const char* s = std::getenv("EK_RDV_FILE"))
    path = s;
  else {
    // one file per LAUNCH: the ranks of one launch share their parent (torchrun's agent, mpirun, the shell), so its
    // pid keys the file together with the rendezvous port and torchrun's run id.  A file left behind by an earlier
    // launch therefore has another name, and rank 0 removes its own after the communicator is up (rendezvous_done).
    const char* port = std::getenv("MASTER_PORT");
    const char* run_id = std::getenv("TORCHELASTIC_RUN_ID");
    path = std::string("/tmp/ek_b200_rdv_") + (port ? port : "0") + "_" + (run_id ? run_id : "none") + "_" +
           std::to_string((long)getppid());
  }
  g_rdv_path = path;
  if (rank == 0) {
    std::remove(path.c_str()); // never publish on top of a stale file
    if (ek_comm_unique_id(id128) != EK_OK) {
      fprintf(stderr, "[euler_b200] NCCL is not available: cannot run on %d ranks\n", nranks);
      std::exit(EXIT_FAILURE);
    }
    const std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id128, 1, EK_COMM_ID_BYTES, f) != EK_COMM_ID_BYTES) {
      fprintf(stderr, "[euler_b200] cannot write rendezvous file %s\n", tmp.c_str());
      std::exit(EXIT_FAILURE);
    }
    std::fclose(f);
    std::rename(tmp.c_str(), path.c_str()); // atomic publish
    return;
  }
  for (int tries = 0; tries < 6000; ++tries) { // up to 10 minutes
    FILE* f = std::fopen(path.c_str(), "rb");
    if (f) {
      const size_t n = std::fread(id128, 1, EK_COMM_ID_BYTES, f);
      std::fclose(f);
      if (n == EK_COMM_ID_BYTES)
        return;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  fprintf(stderr, "[euler_b200] rank %d: timed out waiting for %s\n", rank, path.c_str());
  std::exit(EXIT_FAILURE);
}

} // namespace euler_b200
