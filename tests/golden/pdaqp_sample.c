#include "pdaqp_sample.h"
c_float pdaqp_halfplanes[8] = {
(c_float)0.70175466943160769162,
(c_float)0.12280377788109773929,
(c_float)-0.70175466943160769162,
(c_float)3.51820280548571462376,
(c_float)0.70175466943160769162,
(c_float)0.12280377788109773929,
(c_float)-0.70175466943160769162,
(c_float)-3.51820280548571462376,
};
c_float pdaqp_feedbacks[12] = {
(c_float)-0.19946396163899515752,
(c_float)-0.03490525835793702875,
(c_float)0.19946396163899515752,
(c_float)0.00000000000000000000,
(c_float)0.00000000000000001339,
(c_float)0.00000000000000000181,
(c_float)-0.00000000000000001339,
(c_float)0.99999999999999988898,
(c_float)0.00000000000000001339,
(c_float)0.00000000000000000181,
(c_float)-0.00000000000000001339,
(c_float)-0.99999999999999988898,
};
c_int pdaqp_hp_list[5] = {
(c_int)0,
(c_int)2,
(c_int)1,
(c_int)0,
(c_int)1,
};
c_int pdaqp_jump_list[5] = {
(c_int)1,
(c_int)0,
(c_int)1,
(c_int)0,
(c_int)0,
};
void pdaqp_evaluate(c_float* parameter, c_float* solution){
    int i,j,disp;
    int id,next_id;
    c_float val;
    id = 0;
    next_id = id+pdaqp_jump_list[id];
    while(next_id != id){
        // Compute halfplane value
        disp = pdaqp_hp_list[id]*(PDAQP_N_PARAMETER+1);
        for(i=0, val=0; i<PDAQP_N_PARAMETER; i++)
            val += parameter[i] * pdaqp_halfplanes[disp++];
        id = next_id + (val <= pdaqp_halfplanes[disp]);
        next_id = id+pdaqp_jump_list[id];
    }
    // Leaf node reached -> evaluate affine function
    disp = pdaqp_hp_list[id]*(PDAQP_N_PARAMETER+1)*PDAQP_N_SOLUTION;
    for(i=0; i < PDAQP_N_SOLUTION; i++){
        for(j=0, val=0; j < PDAQP_N_PARAMETER; j++)
            val += parameter[j] * pdaqp_feedbacks[disp++];
        val += pdaqp_feedbacks[disp++];
        solution[i] = val;
    }
}
