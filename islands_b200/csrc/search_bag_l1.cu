// search_bag_l1.cu — see search_bag.inc
#define ISL_BAG_ACC ACC_L1
#define ISL_BAG_SUFFIX l1
#include "search_bag.inc"
