// Shim: included by the reference's pipeline_template.h but not used by it.
#pragma once
