// Instantiations of the cluster tile-pass kernel for 64-bit amplitude indices (registers of more than 31 qubits).
#define QCA_PASS3_WIDE 1
#include "qca_pass3_u32.cu"
