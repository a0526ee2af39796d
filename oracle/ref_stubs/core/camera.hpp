// TEST INFRASTRUCTURE ONLY.  Stand-in for /root/reference/include/core/camera.hpp so that the reference's
// src/training/rasterization/rasterizer_autograd.{hpp,cpp} compile UNCHANGED outside the application: they include this
// header but use nothing from it (the gs::Camera class needs glm, OpenImageIO and the whole loader tree).
#pragma once
namespace gs {
    class Camera;
}
