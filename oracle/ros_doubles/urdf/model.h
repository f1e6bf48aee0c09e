// TEST INFRASTRUCTURE -- test double: the reference only calls initParam (local_planner.cpp:330)
#pragma once
#include <string>
namespace urdf { struct Model { bool initParam(const std::string&) { return true; } }; }
