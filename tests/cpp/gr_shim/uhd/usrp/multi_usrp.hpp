#pragma once
namespace uhd { struct time_spec_t { double secs; }; }
