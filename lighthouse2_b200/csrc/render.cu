/* render.cu - placeholder entry points; replaced by the wavefront loop. */
#include "core.h"
using namespace lh2b;
#define NOTYET( name ) { (void)core; SetLastError( name ": not implemented yet" ); return 1; }
extern "C" {
int lh2b_set_target( lh2b_core* core, int, int, int ) NOTYET( "SetTarget" )
int lh2b_setting( lh2b_core* core, const char*, float ) NOTYET( "Setting" )
int lh2b_set_probe_pos( lh2b_core* core, int, int ) NOTYET( "SetProbePos" )
int lh2b_set_textures( lh2b_core* core, const void*, int ) NOTYET( "SetTextures" )
int lh2b_set_materials( lh2b_core* core, const void*, int ) NOTYET( "SetMaterials" )
int lh2b_set_lights( lh2b_core* core, const void*, int, const void*, int, const void*, int, const void*, int ) NOTYET( "SetLights" )
int lh2b_set_sky( lh2b_core* core, const float*, int, int, const float* ) NOTYET( "SetSkyData" )
int lh2b_render( lh2b_core* core, const void*, int, int ) NOTYET( "Render" )
int lh2b_wait_for_render( lh2b_core* core ) NOTYET( "WaitForRender" )
int lh2b_get_stats( lh2b_core* core, void* ) NOTYET( "GetCoreStats" )
int lh2b_read_pixels( lh2b_core* core, float* ) NOTYET( "ReadPixels" )
int lh2b_read_accumulator( lh2b_core* core, float* ) NOTYET( "ReadAccumulator" )
int lh2b_accumulator_device_ptr( lh2b_core* core, void**, int* ) NOTYET( "AccumulatorDevicePtr" )
int lh2b_set_sample_shard( lh2b_core* core, int, int ) NOTYET( "SetSampleShard" )
int lh2b_get_frame_stats( lh2b_core* core, lh2b_frame_stats* ) NOTYET( "GetFrameStats" )
}
