// TEST INFRASTRUCTURE (oracle/_ref build only).
// Empty stand-in for gproshan's viewer/viewer.h: src/che.cpp:4 includes it but uses nothing
// from it, and the real header needs OpenGL/GLEW/GLUT which this image does not have.
#pragma once
