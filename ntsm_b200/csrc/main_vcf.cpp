// ntsmVCF binary: the reference command line (src/ntSeqMatchVCF.cpp) over the C ABI
#include "../../include/ntsm_b200.h"

int main(int argc, char **argv) { return ntsm_vcf_main(argc, argv); }
