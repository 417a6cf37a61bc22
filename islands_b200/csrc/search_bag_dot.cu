// search_bag_dot.cu — see search_bag.inc
#define ISL_BAG_ACC ACC_DOT
#define ISL_BAG_SUFFIX dot
#include "search_bag.inc"
