// empty stand-in for <Windows.h> (MAIN:17); min/max come from pre.h
#pragma once
