#ifndef MSAM_VERSION_H
#define MSAM_VERSION_H
#define MSAM_GIT_COMMIT "oracle-ref-shim"
#endif
