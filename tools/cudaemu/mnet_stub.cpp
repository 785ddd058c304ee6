// The CPU logic checker covers dlb200.cu only; the recurrent entry points exist so that the binding loads.
#define STUB(name) extern "C" __attribute__((visibility("default"))) long long name(...) { return -1; }
STUB(dlb_mnet_create) STUB(dlb_mnet_destroy) STUB(dlb_mnet_param_count) STUB(dlb_mnet_layer_param_offset) STUB(dlb_mnet_set_params)
STUB(dlb_mnet_get_params) STUB(dlb_mnet_act_host) STUB(dlb_mnet_gradient_sum_host) STUB(dlb_mnet_learn_host) STUB(dlb_mnet_set_batch)
STUB(dlb_mnet_learn_resident)
extern "C" __attribute__((visibility("default"))) const char* dlb_mnet_last_error() { return "not in the CPU logic checker"; }
