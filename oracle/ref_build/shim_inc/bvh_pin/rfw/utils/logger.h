// rfw/utils/logger.h prints through a GL-era utility library; the node sources only use these macros on error paths
#pragma once
#define WARNING(...) ((void)0)
#define DEBUG(...) ((void)0)
#define FAILURE(...) ((void)0)
