// ntsmCount -- same command line as the reference binary (src/ntSeqMatchCount.cpp:53-184);
// all the work happens behind the C ABI in libntsm_b200.so.
#include "../../include/ntsm_b200.h"

int main(int argc, char **argv) { return ntsm_main(argc, argv); }
