#include "CReconstruction.h"

#include <stdio.h>

// CReconstruction.cpp:6-20: open the config, initialise the data manager, wire the matcher (radius 2, ws 0.03) and
// the sink (100, 1, 50, 2, 2.5, no duplicate deletion).
bool CReconstrction::Init(char* configfile) {
  sbcv::FileStorage fs(configfile, sbcv::FileStorage::READ);
  if (fs.isOpened() == false) {
    printf("cannot open file %s\n", configfile);
    return false;
  }
  fs["filepath"] >> filepath;
  if (m_ImageData.Init(fs) == false) return false;
  m_Matching.Init(&m_ImageData, &m_CloudOptimization, 2, 0.03);
  m_CloudOptimization.Init(100, 1, 50, 2, 2.5, &m_ImageData, false);
  return true;
}
