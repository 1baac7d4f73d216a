/* Test-infrastructure shim: the reference's Core/scene_parser.h:20 includes <Windows.h>
 * only for the directory scan below; nothing from it is needed on Linux. */
#pragma once
