#define DEB_HOST_MULTIARCH "x86_64-linux-gnu"
