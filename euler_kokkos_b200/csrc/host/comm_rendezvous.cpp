#include "comm_rendezvous.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#include "../../../include/ek_b200.h"

namespace euler_b200 {

void rendezvous_unique_id(int rank, int nranks, unsigned char* id128)
{
  std::string path;
  if (const char* s = std::getenv("EK_RDV_FILE"))
    path = s;
  else {
    const char* port = std::getenv("MASTER_PORT");
    path = std::string("/tmp/ek_b200_rdv_") + (port ? port : "0");
  }
  if (rank == 0) {
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
