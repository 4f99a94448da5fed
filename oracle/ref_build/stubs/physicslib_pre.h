// Force-included in front of PhysicsLib/convexhull/graham.cpp ONLY: graham.cpp includes ../physicsLib.h, which drags in MainLib
// (FLTK / Ogre, absent here) although the Graham scan needs nothing but BaseLib's matrixn for get_hull.  Defining the include
// guard of physicsLib.h skips it; the BaseLib umbrella header is included instead.
#pragma once
#define STDAFX_OGREFLTK_H
#include "stdafx.h"
#include "math/mathclass.h"
#include <cstdlib>
