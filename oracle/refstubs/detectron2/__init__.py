"""Import-time stand-in so /root/reference imports on a box without detectron2.
Test infrastructure only (golden generation); never imported by the product."""
