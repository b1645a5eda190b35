// Source-compatibility stub: the reference's examplemain.cpp includes "RLBotClient.h" (RLBot runtime glue, out of scope: no GPU work)
// without using it in main().
#pragma once
