// hikari_settings_convert.hpp — hikari_settings (C view, include/hikari_host.h) -> hikari::HikariSettings; shared by the two
// extern "C" shim files.
#pragma once
#include "hikari.hpp"
#include "hikari_host.h"

namespace hikari {
inline HikariSettings to_cpp(const hikari_settings* s) {
    HikariSettings r;
    r.direct_validate_interval = s->direct_validate_interval;
    r.emissive_validate_interval = s->emissive_validate_interval;
    r.max_temporal_reuse_count = s->max_temporal_reuse_count;
    r.max_spatial_reuse_count = s->max_spatial_reuse_count;
    r.max_reservoir_lifetime = s->max_reservoir_lifetime;
    r.solar_angle = s->solar_angle;
    r.indirect_bounces = s->indirect_bounces;
    r.max_indirect_luminance = s->max_indirect_luminance;
    for (int i = 0; i < 4; ++i) r.clear_color[i] = s->clear_color[i];
    r.temporal_reuse = s->temporal_reuse != 0;
    r.emissive_spatial_reuse = s->emissive_spatial_reuse != 0;
    r.indirect_spatial_reuse = s->indirect_spatial_reuse != 0;
    r.denoise = s->denoise != 0;
    r.taa = s->taa == HIKARI_TAA_NONE ? Taa::None : Taa::Jasmine;
    r.upscale.kind = s->upscale_kind == HIKARI_UPSCALE_FSR1 ? Upscale::Fsr1 : Upscale::SmaaTu4x;
    r.upscale.ratio_value = s->upscale_ratio;
    r.upscale.sharpness_value = s->upscale_sharpness;
    return r;
}
}  // namespace hikari
