// TEST INFRASTRUCTURE ONLY (oracle/): plain-C driver around the UNMODIFIED reference biome-map producer
// (SuperTerrain+/SuperTerrain+/Private/World/Diversity/STPBiomeFactory.cpp:24-42, STPLayer.cpp and the demo's layer chain
// SuperDemo+/World/Layers/STPAllLayers.cpp:61-109), compiled from /root/reference by oracle/Makefile. Nothing in the
// product path links or loads this.
#include <World/Layers/STPAllLayers.h>
#include <World/Biomes/STPBiomeRegistry.h>

#include <cstdint>
#include <exception>
#include <memory>

namespace Reg = STPDemo::STPBiomeRegistry;

extern "C" {

// ids: Ocean, Plains, Forest, FrozenOcean, WarmOcean, LukewarmOcean, ColdOcean (SuperDemo+/Biome.ini gives Ocean = 0,
// Plains = 1, Forest = 3; biomes the file does not list keep the zero of their static storage)
void ref_biome_set_ids(const uint16_t ids[7]) {
    Reg::Ocean.ID = ids[0];
    Reg::Plains.ID = ids[1];
    Reg::Forest.ID = ids[2];
    Reg::FrozenOcean.ID = ids[3];
    Reg::WarmOcean.ID = ids[4];
    Reg::LukewarmOcean.ID = ids[5];
    Reg::ColdOcean.ID = ids[6];
}

void* ref_biome_create(uint32_t width, uint32_t height, uint64_t seed) {
    try {
        return new STPDemo::STPLayerChainBuilder(glm::uvec2(width, height), seed);
    } catch (const std::exception&) {
        return nullptr;
    }
}

void ref_biome_destroy(void* factory) { delete static_cast<STPDemo::STPLayerChainBuilder*>(factory); }

int ref_biome_run(void* factory, uint16_t* biomemap, int32_t offset_x, int32_t offset_z) {
    try {
        (*static_cast<STPDemo::STPLayerChainBuilder*>(factory))(biomemap, glm::ivec2(offset_x, offset_z));
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

}  // extern "C"
