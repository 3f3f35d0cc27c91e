// Shared by comm.cu and its kernels: entry ranges of the exchange plan, passed to kernels by value.
#pragma once
#include "lpic_common.cuh"

struct PeerRanges {
    i64 send_first[LPIC_MAX_PEERS + 1], recv_first[LPIC_MAX_PEERS + 1];
};
