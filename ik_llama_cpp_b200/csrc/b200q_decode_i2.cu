// instantiation unit 2 of the decode mat-vec kernels (see b200q_decode_inst.inc)
#define B200Q_INST_GROUP 2
#include "b200q_decode_inst.inc"
