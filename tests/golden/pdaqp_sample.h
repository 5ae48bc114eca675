#ifndef PDAQP_H
#define PDAQP_H

typedef double c_float;
typedef unsigned short c_int;
#define PDAQP_N_PARAMETER 3
#define PDAQP_N_SOLUTION 1

extern c_float pdaqp_halfplanes[8];
extern c_float pdaqp_feedbacks[12];
extern c_int pdaqp_hp_list[5];
extern c_int pdaqp_jump_list[5];
void pdaqp_evaluate(c_float* parameter, c_float* solution);
#endif // ifndef PDAQP_H
