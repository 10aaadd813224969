/* core_api_filter.cpp - libRenderCore_B200Filter.so: the drop-in for the reference's RenderCore_Optix7Filter
   (lib/RenderCore_Optix7Filter/core_api.cpp:18-22 exports the same CreateCore). A forwarding library: it links against
   libRenderCore_B200.so next to it ($ORIGIN) and returns that library's filter persona, which honours Setting("filter"),
   ("TAA"), ("clampDirect"), ("clampIndirect") (see ../core_api.cpp). */
namespace lh2abi { class CoreAPI_Base; }
extern "C" lh2abi::CoreAPI_Base* CreateCoreFilter();
extern "C" __attribute__( ( visibility( "default" ) ) ) lh2abi::CoreAPI_Base* CreateCore() { return CreateCoreFilter(); }
