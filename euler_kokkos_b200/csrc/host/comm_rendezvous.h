// File-based rendezvous of the NCCL unique id between the ranks of a torchrun-style launch
// (one process per GPU).  Stands in for the reference's MPI_Init / MPI_COMM_WORLD bootstrap
// (src/utils/mpi/GlobalMpiSession.cpp); payload is the 128-byte id only.
#pragma once

namespace euler_b200 {
//! rank 0 creates the id (ek_comm_unique_id) and publishes it; the others wait for it.
//! The file is $EK_RDV_FILE or /tmp/ek_b200_rdv_<MASTER_PORT|0>_<TORCHELASTIC_RUN_ID|none>_<pid of the launcher>:
//! one name per launch, so a file left by an earlier launch on the same port is never read.
void rendezvous_unique_id(int rank, int nranks, unsigned char* id128);
//! after the communicator exists on every rank (ek_comm_attach returned): rank 0 removes the file.
void rendezvous_done(int rank);
} // namespace euler_b200
