// TEST INFRASTRUCTURE ONLY.  Stand-in for /root/reference/include/core/splat_data.hpp (see camera.hpp beside it): the
// reference's rasterizer_autograd.{hpp,cpp} include it but use nothing from it.
#pragma once
namespace gs {
    class SplatData;
}
