// search_bag_l2.cu — see search_bag.inc
#define ISL_BAG_ACC ACC_L2
#define ISL_BAG_SUFFIX l2
#include "search_bag.inc"
