// File-based rendezvous of the NCCL unique id between the ranks of a torchrun-style launch
// (one process per GPU).  Stands in for the reference's MPI_Init / MPI_COMM_WORLD bootstrap
// (src/utils/mpi/GlobalMpiSession.cpp); payload is the 128-byte id only.
#pragma once

namespace euler_b200 {
//! rank 0 creates the id (ek_comm_unique_id) and publishes it; the others wait for it.
//! The file is $EK_RDV_FILE or /tmp/ek_b200_rdv_<MASTER_PORT|0>.
void rendezvous_unique_id(int rank, int nranks, unsigned char* id128);
} // namespace euler_b200
