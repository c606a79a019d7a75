// CReconstruction.h — host mirror of the reference's top-level class (reconstruction/CReconstruction.h:12-20).
#pragma once
#include "CStereoMatching.h"

#define USAGE_HELP "help me"

class CReconstrction {  // spelling as in the reference
 public:
  CManageData m_ImageData;
  CStereoMatching m_Matching;
  CCloudOptimization m_CloudOptimization;
  std::string filepath;
  bool Init(char* configfile);
};
